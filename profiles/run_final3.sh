mkdir -p gpurun_out/final3 gpurun_out/prof3; O=gpurun_out/final3; P=gpurun_out/prof3
timeout 400 python bench.py > $O/bench_default.json 2> $O/bench_default.err
timeout 200 python bench.py --batch 64 --no-extra --no-cpu-baseline --all-kernels > $O/bench_64.json 2> $O/bench_64.err
timeout 200 python bench.py --no-extra --no-cpu-baseline --all-kernels > $O/bench_1024_all.json 2> $O/bench_1024_all.err
timeout 200 python bench.py --precision tf32 --no-extra --no-cpu-baseline > $O/bench_tf32.json 2> $O/bench_tf32.err
timeout 200 python bench.py --precision tf32x3c --no-extra --no-cpu-baseline > $O/bench_tf32x3c.json 2> $O/bench_tf32x3c.err
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; tail -1 $O/smoke.log
BENCH="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extra --no-profile"
AVA_B200_SIDE_STREAM=0 timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file $P/r02_launches.csv $BENCH > $P/ncu_launches.log 2>&1
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__inst_issued.avg.pct_of_peak_sustained_active,sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__grid_size,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio
AVA_B200_SIDE_STREAM=0 timeout 600 ncu --metrics $M --clock-control none -k regex:"gconv|wgrad|finalize|tc_gemm|tc_split|sgemm|splitk|border|recon|adam|latent|channel_stats|bn_|colsum|col_sum|elbo|bias" -s 200 -c 560 -o $P/r02_step $BENCH > $P/ncu_step.log 2>&1
ncu -i $P/r02_step.ncu-rep --page raw --csv > $P/r02_step_raw.csv 2>/dev/null
rm -f $P/r02_step.ncu-rep
python - <<PY
import json
for f in ("bench_default","bench_64"):
  d=json.loads(open("gpurun_out/final3/%s.json"%f).read().strip().splitlines()[-1])
  print(f,d["ms_per_step"],d["value"],d["e2e"]["value"],d["roofline"]["frac"],[(k["call"],k["us_per_step"]) for k in d["kernel_families"][:3]])
PY
