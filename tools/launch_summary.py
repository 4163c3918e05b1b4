"""Per-kernel summary of an ncu launch list (--csv with gpu__time_duration.sum and the lane / issue / occupancy / DRAM
metrics): total time, share, launches.  Usage: python tools/launch_summary.py profiles/<list>.csv"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = next(r for r in rows if "Kernel Name" in r)
ki, mi, vi, idi = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("ID")
per = {}
for r in rows[rows.index(hdr) + 1:]:
    if len(r) <= vi:
        continue
    name = r[ki].split("(")[0].replace("void ", "").replace("vp::", "")[:28]
    per.setdefault((int(r[idi]), name), {})[r[mi]] = float(r[vi].replace(",", ""))
agg = {}
for (_, name), m in per.items():
    a = agg.setdefault(name, dict(n=0, us=0.0, lanes=0.0, issue=0.0, warps=0.0, dram=0.0))
    us = m.get("gpu__time_duration.sum", 0.0) / 1e3
    a["n"] += 1; a["us"] += us
    a["lanes"] += us * m.get("smsp__thread_inst_executed_per_inst_executed.ratio", 0.0)
    a["issue"] += us * m.get("smsp__issue_active.avg.pct_of_peak_sustained_active", 0.0)
    a["warps"] += us * m.get("sm__warps_active.avg.pct_of_peak_sustained_active", 0.0)
    a["dram"] += m.get("dram__bytes_read.sum", 0.0) + m.get("dram__bytes_write.sum", 0.0)
tot = sum(a["us"] for a in agg.values())
print(f"# per-kernel summary of {sys.argv[1]} (serialised, cold-cache launch times: compare SHARES, not absolutes)")
print(f"{'kernel':30s} {'launches':>8s} {'total_us':>10s} {'share':>7s} {'thr/inst':>9s} {'issue%':>7s} {'warps%':>7s} {'dram_MB/launch':>15s}")
for name, a in sorted(agg.items(), key=lambda kv: -kv[1]["us"]):
    w = max(a["us"], 1e-9)
    print(f"{name:30s} {a['n']:8d} {a['us']:10.1f} {100 * a['us'] / tot:6.1f}% {a['lanes'] / w:9.1f} {a['issue'] / w:7.1f} {a['warps'] / w:7.1f} {a['dram'] / a['n'] / 1e6:15.2f}")
