#!/usr/bin/env python3
"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum[,dram__bytes_*] --csv` launch list."""
import collections
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = rows[0]
ki, mi, vi = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value")
agg = collections.defaultdict(lambda: [0, 0.0, 0.0])
for r in rows[1:]:
    k = r[ki].replace("b2rt::", "").replace("<unnamed>::", "").replace("unnamed>::", "").replace("void ", "").split("(")[0][:44]
    v = float(r[vi].replace(",", ""))
    if r[mi] == "gpu__time_duration.sum":
        agg[k][0] += 1; agg[k][1] += v
    elif r[mi].startswith("dram__bytes"):
        agg[k][2] += v
tot = sum(v[1] for v in agg.values())
print(f"{'kernel':44s} {'n':>5s} {'total us':>10s} {'share':>6s} {'DRAM MB':>10s} {'GB/s':>7s}")
for k, v in sorted(agg.items(), key=lambda x: -x[1][1]):
    print(f"{k:44s} {v[0]:5d} {v[1] / 1e3:10.1f} {v[1] / tot * 100:5.1f}% {v[2] / 1e6:10.1f} {v[2] / max(v[1], 1):7.0f}")
print(f"{'total':44s} {sum(v[0] for v in agg.values()):5d} {tot / 1e3:10.1f}")
