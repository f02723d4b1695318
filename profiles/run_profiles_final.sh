# Round-2 final profile set (run on the GPU box: gpurun -- 'bash profiles/run_profiles_final.sh').
# 1. launch list of two eager steps (per-kernel durations; cold-cache, serialised: shares, not absolutes)
# 2. per-kernel counter table of one eager step (DRAM bytes, pipe utilisation, stalls)
# 3. ncu --set full (+source) of the top kernel of each family
set -x
OUT=gpurun_out/prof
mkdir -p $OUT
BENCH="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extra --no-profile"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file $OUT/r02_launches.csv $BENCH > $OUT/ncu_launches.log 2>&1
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__inst_issued.avg.pct_of_peak_sustained_active,sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__grid_size,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio
timeout 900 ncu --metrics $M --clock-control none -k regex:"gconv|wgrad|finalize|tc_gemm|tc_split|sgemm|splitk|border|recon|adam|latent|channel_stats|bn_|colsum|col_sum|elbo" \
    -s 200 -c 560 -o $OUT/r02_step $BENCH > $OUT/ncu_step.log 2>&1
ncu -i $OUT/r02_step.ncu-rep --page raw --csv > $OUT/r02_step_raw.csv 2>/dev/null
rm -f $OUT/r02_step.ncu-rep
i=0
for RX in 'wgrad_mma_kernel<\(int\)1, \(int\)24, \(int\)16, \(int\)32, \(int\)0, \(int\)2' \
          'gconv_kernel<\(int\)0, \(int\)16, \(int\)24, \(int\)32, \(int\)0, \(int\)0' \
          'gconv_kernel<\(int\)0, \(int\)24, \(int\)16, \(int\)32, \(int\)1, \(int\)1' \
          'gconv_kernel<\(int\)2, \(int\)8, \(int\)8, \(int\)32, \(int\)1, \(int\)1' \
          'tc_gemm_kernel<\(int\)3'; do
  i=$((i+1))
  timeout 300 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"$RX" \
      -c 1 -o $OUT/r02_top$i $BENCH > $OUT/ncu_top$i.log 2>&1
  ncu -i $OUT/r02_top$i.ncu-rep --page raw --csv > $OUT/r02_top${i}_raw.csv 2>/dev/null
  ncu -i $OUT/r02_top$i.ncu-rep --page source --csv --print-source sass > $OUT/r02_top${i}_source.csv 2>/dev/null
  rm -f $OUT/r02_top$i.ncu-rep
done
du -sh $OUT
