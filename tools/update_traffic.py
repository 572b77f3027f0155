#!/usr/bin/env python3
"""profiles/traverse_traffic.json (read by bench.py for roofline.traffic) from an ncu launch list that holds
dram__bytes_read.sum / dram__bytes_write.sum per launch:   python tools/update_traffic.py profiles/launches_r02_cfg3.csv cfg3
The file is keyed by bench.py workload (cfg2, cfg3, ...)."""
import csv
import json
import os
import sys

src = sys.argv[1]
workload = sys.argv[2] if len(sys.argv) > 2 else "cfg3"
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
                 "one frame of `bench.py --workload ` + workload + ` --steps 1 --warmup 0 --no-cpu`-class run named in the file header, all traversal launches)",
       "note": "dram__bytes_read.sum + dram__bytes_write.sum averaged over the traversal launches of the frame"}
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
path = os.path.join(root, "profiles", "traverse_traffic.json")
allw = json.load(open(path)) if os.path.exists(path) else {}
if "dram_bytes_per_launch" in allw:      # round-1 layout (one workload): start over
    allw = {}
allw[workload] = out
json.dump(allw, open(path, "w"), indent=1)
print(out)
