"""Summarise an `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv` launch list:
per kernel (template arguments kept) launches, share of the captured time, average duration, DRAM bytes per launch,
and the same aggregated per C-ABI entry point (kernel family)."""
import csv, json, re, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hdr = next(r for r in rows if "Kernel Name" in r)
rows = rows[rows.index(hdr) + 1:]
iN, iM, iV = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value")
iID = hdr.index("ID")
per = collections.OrderedDict()
for r in rows:
    if len(r) <= iV: continue
    d = per.setdefault(r[iID], {"name": r[iN]})
    d[r[iM]] = float(r[iV].replace(",", ""))
def short(n):
    n = re.sub(r"^void ", "", n); n = re.sub(r"ava::", "", n); n = re.sub(r"\(int\)", "", n)
    return n.split("(")[0] if "<" not in n else n[:n.index(">") + 1]
def family(n):
    m = re.search(r"gconv_kernel<(\d+), (\d+), (\d+), (\d+), (\d+), (\d+)", n)
    if m: return "bnconv_fwd" if m.group(6) == "0" else "bnconv_bwd_data"
    if "wgrad" in n or "reduce_partials" in n: return "bnconv_bwd_weight"
    if "bn_relu_bwd_apply" in n: return "bn_relu_bwd_apply"
    if any(k in n for k in ("tc_gemm", "tc_split", "sgemm", "splitk_reduce")): return "linear"
    if "adam" in n: return "adam_step"
    if "recon" in n: return "recon"
    if "channel_stats" in n: return "channel_stats"
    return "other"
kern, fam = collections.OrderedDict(), collections.OrderedDict()
tot = 0.0
for d in per.values():
    t = d.get("gpu__time_duration.sum", 0.0) / 1e3   # ns -> us
    b = d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0)
    tot += t
    for table, key in ((kern, short(d["name"])), (fam, family(short(d["name"])))):
        e = table.setdefault(key, {"launches": 0, "us": 0.0, "dram_bytes": 0.0})
        e["launches"] += 1; e["us"] += t; e["dram_bytes"] += b
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
out = {"captured_us": tot, "steps": steps, "families": {}, "kernels": {}}
for table, name in ((fam, "families"), (kern, "kernels")):
    for k, e in sorted(table.items(), key=lambda kv: -kv[1]["us"]):
        out[name][k] = {"launches": e["launches"], "share": round(e["us"] / tot, 4), "avg_us": round(e["us"] / e["launches"], 2),
                        "us_per_step": round(e["us"] / steps, 1), "dram_bytes_per_launch": round(e["dram_bytes"] / e["launches"])}
json.dump(out, open(sys.argv[3], "w"), indent=1) if len(sys.argv) > 3 else None
print("captured %.1f us over %d steps" % (tot, steps))
for k, e in out["families"].items(): print("%-22s" % k, e)
for k, e in list(out["kernels"].items())[:int(sys.argv[4]) if len(sys.argv) > 4 else 16]: print("%-60s" % k[:60], e["launches"], e["share"], e["avg_us"])
