"""Print a compact set of metrics per kernel from an .ncu-rep (raw page)."""
import csv, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[0]
keys = {
 "t_us": "gpu__time_duration.sum", "dramR": "dram__bytes_read.sum", "dramW": "dram__bytes_write.sum",
 "dram%": "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
 "issue%": "smsp__issue_active.avg.pct_of_peak_sustained_active",
 "fma%": "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
 "tensor%": "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
 "lsu%": "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
 "smemWave%": "l1tex__data_pipe_lsu_wavefronts_mem_shared.avg.pct_of_peak_sustained_elapsed",
 "warps%": "sm__warps_active.avg.pct_of_peak_sustained_active", "regs": "launch__registers_per_thread",
 "bankconf": "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
 "st_barrier": "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
 "st_short": "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
 "st_long": "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
 "st_mio": "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
 "st_math": "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
 "st_wait": "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
 "st_notsel": "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
 "st_disp": "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
 "st_lg": "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
 "st_tex": "smsp__average_warps_issue_stalled_tex_throttle_per_issue_active.ratio",
}
ni = hdr.index("Kernel Name")
seen = set()
for r in rows[2:]:
    name = r[ni][:70]
    if name in seen and "--all" not in sys.argv: continue
    seen.add(name)
    vals = []
    for k, m in keys.items():
        if m in hdr:
            v = r[hdr.index(m)]
            try: v = "%.3g" % float(v.replace(",", ""))
            except ValueError: pass
            vals.append("%s=%s" % (k, v))
    print(name); print("   " + " ".join(vals))
