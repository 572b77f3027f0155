#!/usr/bin/env python3
"""Register / spill / shared-memory table of every kernel from the ptxas logs the Makefile writes (csrc/*.o.log):
  python tools/register_table.py > profiles/r02_registers.txt"""
import glob
import os
import re
import subprocess

root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rows = []
for log in sorted(glob.glob(os.path.join(root, "cuda-raytracer_b200", "csrc", "*.o.log"))):
    name = None
    spill = ""
    for ln in open(log):
        m = re.search(r"Compiling entry function '(\S+)' for 'sm_100a'", ln)
        if m:
            name = subprocess.check_output(["cu++filt", m.group(1)], text=True).strip()
            name = name.replace("(int)", "").replace("(bool)", "").replace("void ", "")
            name = re.sub(r"\(.*$", "", name).replace("b2rt::", "").replace("(anonymous namespace)::", "").replace("<unnamed>::", "")
            continue
        m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", ln)
        if m:
            spill = f"{m.group(2)}/{m.group(3)}"
            continue
        m = re.search(r"Used (\d+) registers(?:, used (\d+) barriers)?(?:, (\d+) bytes cumulative stack size)?(?:, (\d+) bytes smem)?", ln)
        if m and name:
            smem = re.search(r"(\d+) bytes smem", ln)
            rows.append((os.path.basename(log)[:-6], name, int(m.group(1)), spill, smem.group(1) if smem else "0"))
            name = None
print("ptxas -v, nvcc 12.9, -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false (csrc/Makefile); spills = store/load bytes")
print(f"{'object':16} {'kernel':64} {'regs':>5} {'spills':>9} {'static smem':>12}")
for r in rows:
    print(f"{r[0]:16} {r[1][:64]:64} {r[2]:5d} {r[3]:>9} {r[4]:>12}")
