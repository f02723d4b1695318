#!/bin/bash
# Round-2 profiling recipe (run under gpurun from the repo root; writes only small files to gpurun_out/).
#   1. launch list of the bench command (per-launch device time; compare SHARES, not absolutes)
#   2. per-kernel counters of one eager train step (DRAM bytes, pipe utilisation, occupancy, stalls)
#   3. `--set full --import-source on` captures of the top kernels and of get_spec
#   4. the two micro-probes behind DESIGN.md 3.1
set -x
OUT=gpurun_out
BENCH="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra --no-profile"
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $OUT/r02_launches.csv $BENCH > $OUT/ncu_bench.log 2>&1
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__inst_issued.avg.pct_of_peak_sustained_active,sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__grid_size,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio
# one eager step = launches 2 of the bench (the first two steps at a batch size run eagerly): skip step 1
timeout 900 ncu --metrics $M --clock-control none -k regex:"gconv|wgrad|finalize|tc_gemm|tc_split|sgemm|splitk|border|recon|adam|latent|channel_stats|bn_" \
    -s 170 -c 175 -o $OUT/r02_step python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extra --no-profile > $OUT/ncu_step.log 2>&1
ncu -i $OUT/r02_step.ncu-rep --page raw --csv > $OUT/r02_step_raw.csv 2>/dev/null
ls -la $OUT/r02_step.ncu-rep; rm -f $OUT/r02_step.ncu-rep
# full-set captures of the top kernels (a handful of launches each)
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"wgrad_mma_kernel<1, 24, 16|wgrad_mma_kernel<1, 16, 8|gconv_kernel<0, 16, 8, 32, 0, 0|gconv_kernel<2, 8, 8, 32, 1, 1|gconv_kernel<0, 16, 24, 32, 1, 1|tc_gemm_kernel<3" \
    -s 12 -c 8 -o $OUT/r02_top python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extra --no-profile > $OUT/ncu_top.log 2>&1
ncu -i $OUT/r02_top.ncu-rep --page raw --csv > $OUT/r02_top_raw.csv 2>/dev/null
ls -la $OUT/r02_top.ncu-rep
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"get_spec" -s 2 -c 1 -o $OUT/r02_spec python profiles/probes/spec_only.py > $OUT/ncu_spec.log 2>&1
ncu -i $OUT/r02_spec.ncu-rep --page raw --csv > $OUT/r02_spec_raw.csv 2>/dev/null
ncu -i $OUT/r02_spec.ncu-rep --page source --csv > $OUT/r02_spec_source.csv 2>/dev/null
ls -la $OUT/r02_spec.ncu-rep
timeout 120 ./profiles/probes/hmma_lat > $OUT/probe_hmma.txt 2>&1
timeout 120 ./profiles/probes/tc_probe > $OUT/probe_tc.txt 2>&1
du -sh $OUT
# keep the merged directory under the 64 MiB limit
for f in $OUT/r02_top.ncu-rep $OUT/r02_spec.ncu-rep; do
  if [ -f $f ] && [ $(stat -c %s $f) -gt 20000000 ]; then rm -f $f; fi
done
du -sh $OUT
