# ncu --set full (+source) of the top kernels of one eager step; run on the GPU box:
#   gpurun -- 'bash profiles/run_profiles_top.sh'
set -x
OUT=gpurun_out
RX='gconv_kernel<\(int\)0, \(int\)16, \(int\)24, \(int\)32, \(int\)0, \(int\)0|wgrad_mma_kernel<\(int\)1, \(int\)24, \(int\)16, \(int\)32, \(int\)0|gconv_kernel<\(int\)2, \(int\)8, \(int\)8, \(int\)32, \(int\)1, \(int\)1|tc_gemm_kernel<\(int\)3'
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"$RX" \
    -c 6 -o $OUT/r02_top python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extra --no-profile > $OUT/ncu_top.log 2>&1
ncu -i $OUT/r02_top.ncu-rep --page raw --csv > $OUT/r02_top_raw.csv 2>/dev/null
ncu -i $OUT/r02_top.ncu-rep --page source --csv --print-source sass > $OUT/r02_top_source.csv 2>/dev/null
ls -la $OUT/r02_top.ncu-rep
