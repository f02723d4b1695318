#!/usr/bin/env python
"""Compact report of `ncu --set full --import-source on` captures exported on the GPU box as
<name>_raw.csv (--page raw) and <name>_source.csv (--page source --print-source sass):
key counters, stall-reason samples, SASS instruction mix, and the hottest SASS lines.

  python profiles/summarize_ncu_full.py gpurun_out/prof/r02_top1 [more prefixes] > profiles/r02_ncu_top_kernels.txt
"""
import csv
import io
import re
import sys
from collections import Counter

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"), ("launch__block_size", "block"),
    ("launch__registers_per_thread", "regs/thread"),
    ("launch__shared_mem_per_block_dynamic", "dyn smem/block"),
    ("launch__occupancy_limit_registers", "occupancy limit (regs), blocks/SM"),
    ("launch__occupancy_limit_shared_mem", "occupancy limit (smem), blocks/SM"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
    ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput %"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate %"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1/shared throughput %"),
    ("sm__inst_issued.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active %"),
    ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "FMA pipe active %"),
    ("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "ALU pipe active %"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU pipe %"),
    ("sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active", "uniform pipe %"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "shared bank conflicts"),
    ("smsp__inst_executed.sum", "warp instructions executed"),
]


def read_raw(path):
    raw = open(path).read()
    rd = csv.reader(io.StringIO(raw[raw.find('"ID"'):]))
    h, u, r = next(rd), next(rd), next(rd)
    return {a: (c, b) for a, b, c in zip(h, u, r)}


def report(prefix):
    m = read_raw(prefix + "_raw.csv")
    print("=" * 110)
    print(m["Kernel Name"][0])
    for key, label in KEYS:
        if key in m and m[key][0] != "":
            print("  %-44s %s %s" % (label, m[key][0], m[key][1]))
    stalls = []
    for k, (v, _) in m.items():
        mm = re.match(r"smsp__pcsamp_warps_issue_stalled_(\w+)$", k)
        if mm and not mm.group(1).endswith("not_issued") and v not in ("", "0"):
            stalls.append((float(v), mm.group(1)))
    tot = sum(v for v, _ in stalls) or 1.0
    print("  warp-state samples: " + ", ".join("%s %.0f%%" % (n, 100 * v / tot) for v, n in sorted(stalls, reverse=True)[:8]))
    # source page
    raw = open(prefix + "_source.csv").read()
    rd = csv.reader(io.StringIO(raw[raw.find('"Address"'):]))
    h = next(rd)
    col = {n: i for i, n in enumerate(h)}
    rows = [r for r in rd if len(r) >= len(h) - 2]
    mix = Counter()
    execd = Counter()
    for r in rows:
        op = r[col["Source"]].strip()
        op = re.sub(r"^@!?U?P\d+\s+", "", op)
        mnem = op.split()[0] if op else "?"
        base = mnem.split(".")[0]
        try:
            n = float(r[col["Instructions Executed"]] or 0)
        except ValueError:
            n = 0.0
        mix[base] += 1
        execd[mnem if base in ("HMMA", "UTCHMMA", "UTMALDG", "LDTM", "UTCBAR", "SYNCS", "F2FP") else base] += n
    te = sum(execd.values()) or 1.0
    print("  executed warp-instruction mix: " + ", ".join("%s %.1f%%" % (k, 100 * v / te) for k, v in execd.most_common(12)))
    def samples(r):
        try:
            return float(r[col["# Samples"]] or 0)
        except ValueError:
            return 0.0
    ts = sum(samples(r) for r in rows) or 1.0
    stall_cols = [n for n in h if n.startswith("stall_") and "Not Issued" not in n]
    print("  hottest SASS lines (share of all warp-state samples; dominant states):")
    for r in sorted(rows, key=samples, reverse=True)[:10]:
        st = []
        for n in stall_cols:
            try:
                v = float(r[col[n]] or 0)
            except ValueError:
                v = 0.0
            if v > 0:
                st.append((v, n[6:]))
        st = ", ".join("%s %.0f" % (n, v) for v, n in sorted(st, reverse=True)[:3])
        print("    %5.1f%%  %-64s %s" % (100 * samples(r) / ts, r[col["Source"]].strip()[:64], st))


if __name__ == "__main__":
    for p in sys.argv[1:]:
        report(p)
