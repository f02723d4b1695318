#!/usr/bin/env python
"""Summarise ncu output for profiles/ (run HERE, on the CPU box, on files brought back in gpurun_out/).

  python profiles/summarize_ncu.py launches gpurun_out/launches.csv  > profiles/rNN_ncu_launch_summary.json
      launch list of `ncu --metrics gpu__time_duration.sum --clock-control none --csv`: per kernel
      family (the C-ABI entry point it belongs to) launches, total time and share of the step.

  python profiles/summarize_ncu.py full gpurun_out/prof.ncu-rep      > profiles/rNN_ncu_top_kernels.txt
      (also writes profiles/rNN_dram_traffic.json when given --traffic FILE)
      one `ncu --set full` capture of an eager step: per kernel duration, DRAM bytes, pipe / issue
      utilisation, occupancy, registers, the main stall reasons.
"""
import csv
import io
import json
import re
import subprocess
import sys
from collections import OrderedDict, defaultdict


def family(kernel):
    """C-ABI entry point a kernel belongs to (csrc/*.cu)."""
    k = kernel
    m = re.search(r"gconv_kernel<(\d+), *(\d+), *(\d+), *(\d+), *(\d+), *(\d+)", k)
    if m:
        return "bnconv_bwd_data" if m.group(6) == "1" else "bnconv_fwd"
    if "wgrad" in k or "bnconv_finalize" in k:
        return "bnconv_bwd_weight"
    if "dz_border_sums" in k:
        return "dz_border_sums"
    if "tc_gemm" in k or "tc_split" in k or "sgemm" in k or "splitk" in k or "colsum" in k or "col_sum" in k:
        return "linear_*"
    if "adam" in k:
        return "adam_step"
    if "recon" in k or "latent" in k or "elbo" in k:
        return "elbo"
    if "channel_stats" in k or "bn_" in k:
        return "bn bookkeeping"
    if "get_spec" in k or "time_tables" in k or "quantile" in k:
        return "get_spec"
    return "other (torch: memset / copy / rng)"


def launches(path):
    rows = []
    with open(path, newline="") as f:
        text = f.read()
    start = text.find('"ID"')
    rd = csv.DictReader(io.StringIO(text[start:]))
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        us = v / 1000.0 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1000.0)
        rows.append((r["Kernel Name"], us))
    fam = defaultdict(lambda: [0, 0.0])
    ker = defaultdict(lambda: [0, 0.0])
    for k, us in rows:
        f = fam[family(k)]
        f[0] += 1
        f[1] += us
        short = re.sub(r"\(.*", "", k)
        kk = ker[short]
        kk[0] += 1
        kk[1] += us
    tot = sum(v[1] for v in fam.values())
    out = {"launches": len(rows), "total_us": round(tot, 1),
           "families": OrderedDict((k, {"launches": v[0], "us": round(v[1], 1), "share": round(v[1] / tot, 4)})
                                   for k, v in sorted(fam.items(), key=lambda kv: -kv[1][1])),
           "kernels": OrderedDict((k, {"launches": v[0], "us": round(v[1], 1), "share": round(v[1] / tot, 4)})
                                  for k, v in sorted(ker.items(), key=lambda kv: -kv[1][1])[:40])}
    print(json.dumps(out, indent=1))


WANT = OrderedDict([
    ("gpu__time_duration.sum", "t_us"),
    ("dram__bytes_read.sum", "dramR_MB"),
    ("dram__bytes_write.sum", "dramW_MB"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
    ("sm__inst_issued.avg.pct_of_peak_sustained_active", "issue%"),
    ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "fma%"),
    ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "fp64%"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor%"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "lsu%"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps%"),
    ("launch__registers_per_thread", "regs"),
    ("launch__grid_size", "grid"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "bankconf"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "st_barrier"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "st_long"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "st_short"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "st_math"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "st_wait"),
    ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "st_mio"),
    ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "st_lg"),
])


def full(path, traffic_path=None):
    if path.endswith(".csv"):      # already exported on the GPU box: ncu -i X.ncu-rep --page raw --csv
        raw = open(path).read()
    else:
        raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    start = raw.find('"ID"')
    rd = csv.reader(io.StringIO(raw[start:]))
    header = next(rd)
    units = next(rd)
    col = {h: i for i, h in enumerate(header)}
    kcol = col["Kernel Name"]
    agg = OrderedDict()
    traffic = defaultdict(float)
    for r in rd:
        if len(r) < len(header):
            continue
        name = r[kcol]
        vals = {}
        for metric, short in WANT.items():
            if metric in col:
                try:
                    v = float(r[col[metric]].replace(",", ""))
                except ValueError:
                    continue
                u = units[col[metric]]
                if short == "t_us":
                    v = v / 1000.0 if u in ("ns", "nsecond") else (v if u in ("us", "usecond") else v * 1000.0)
                if short in ("dramR_MB", "dramW_MB"):
                    scale = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(u, 1e-6)
                    v = v * scale
                vals[short] = v
        traffic[family(name)] += (vals.get("dramR_MB", 0.0) + vals.get("dramW_MB", 0.0)) * 1e6
        a = agg.setdefault(name, {"n": 0})
        a["n"] += 1
        for k, v in vals.items():
            a[k] = a.get(k, 0.0) + v
    for name, a in sorted(agg.items(), key=lambda kv: -kv[1].get("t_us", 0.0)):
        n = a.pop("n")
        t = a.get("t_us", 0.0)
        print(name[:150])
        parts = ["launches=%d" % n, "t_us(total)=%.1f" % t]
        for k in WANT.values():
            if k in ("t_us",) or k not in a:
                continue
            v = a[k] if k in ("dramR_MB", "dramW_MB", "bankconf") else a[k] / n
            parts.append("%s=%.3g" % (k, v))
        print("   " + " ".join(parts))
    if traffic_path:
        with open(traffic_path, "w") as f:
            json.dump({k: int(v) for k, v in traffic.items()}, f, indent=1)


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2])
    else:
        tp = sys.argv[sys.argv.index("--traffic") + 1] if "--traffic" in sys.argv else None
        full(sys.argv[2], tp)
