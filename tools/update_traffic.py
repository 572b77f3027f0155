#!/usr/bin/env python3
"""profiles/traverse_traffic.json (read by bench.py for roofline.traffic) from an ncu launch list that holds
dram__bytes_read.sum / dram__bytes_write.sum per launch:   python tools/update_traffic.py profiles/launches_r01h.csv"""
import csv
import json
import os
import sys

src = sys.argv[1]
rows = [r for r in csv.reader(open(src)) if len(r) > 5]
hdr = rows[0]
ki, mi, ui, vi = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Unit"), hdr.index("Metric Value")
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
tot, n = 0.0, 0
for r in rows[1:]:
    if "k_traverse" not in r[ki]:
        continue
    if r[mi].startswith("dram__bytes"):
        tot += float(r[vi].replace(",", "")) * scale.get(r[ui], 1.0)
    elif r[mi] == "gpu__time_duration.sum":
        n += 1
out = {"kernel": "k_traverse", "dram_bytes_per_launch": tot / max(1, n), "launches": n,
       "source": f"{src} (ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none, "
                 "one frame of the command named in the file header / profiles/r01_launch_summary.txt, all traversal launches)",
       "note": "dram__bytes_read.sum + dram__bytes_write.sum averaged over the traversal launches of the frame"}
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
json.dump(out, open(os.path.join(root, "profiles", "traverse_traffic.json"), "w"), indent=1)
print(out)
