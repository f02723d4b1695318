set -x
OUT=gpurun_out
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__inst_issued.avg.pct_of_peak_sustained_active,sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__grid_size,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio
timeout 900 ncu --metrics $M --clock-control none -k regex:"gconv|wgrad|finalize|tc_gemm|tc_split|sgemm|splitk|border|recon|adam|latent|channel_stats|bn_|colsum|col_sum|elbo" \
    -s 200 -c 560 -o $OUT/r02_step python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extra --no-profile > $OUT/ncu_step.log 2>&1
ncu -i $OUT/r02_step.ncu-rep --page raw --csv > $OUT/r02_step_raw.csv 2>/dev/null
ls -la $OUT/r02_step.ncu-rep; rm -f $OUT/r02_step.ncu-rep
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"wgrad_mma_kernel<1, 24, 16|wgrad_mma_kernel<1, 16, 8|gconv_kernel<0, 16, 8, 32, 0, 0|gconv_kernel<2, 8, 8, 32, 1, 1|gconv_kernel<0, 16, 24, 32, 1, 1|tc_gemm_kernel<3" \
    -s 12 -c 8 -o $OUT/r02_top python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extra --no-profile > $OUT/ncu_top.log 2>&1
ncu -i $OUT/r02_top.ncu-rep --page raw --csv > $OUT/r02_top_raw.csv 2>/dev/null
ls -la $OUT/r02_top.ncu-rep
if [ -f $OUT/r02_top.ncu-rep ] && [ $(stat -c %s $OUT/r02_top.ncu-rep) -gt 30000000 ]; then rm -f $OUT/r02_top.ncu-rep; fi
du -sh $OUT
