#!/usr/bin/env python3
"""Instruction-level dump of an `ncu --set full --import-source on` capture joined with nvdisasm line info:
one row per SASS instruction in program order (executed warp instructions, active threads, stall samples, source line).

  python tools/ncu_sass.py gpurun_out/prof.ncu-rep cuda-raytracer_b200/csrc/traverse.o 'k_traverse<(int)4, (bool)0, (bool)0' [launch#]
"""
import csv
import io
import re
import subprocess
import sys

sys.path.insert(0, __import__("os").path.dirname(__file__))
from ncu_lines import sass_lines  # noqa: E402


def main():
    rep, obj, kre = sys.argv[1], sys.argv[2], sys.argv[3]
    which = int(sys.argv[4]) if len(sys.argv) > 4 else 0
    raw = subprocess.check_output(["ncu", "-i", rep, "--page", "source", "--csv"], text=True, stderr=subprocess.DEVNULL)
    blocks, cur = [], None
    for row in csv.reader(io.StringIO(raw)):
        if row and row[0] == "Kernel Name":
            cur = {"name": row[1], "rows": []}
            blocks.append(cur)
        elif cur is not None and row:
            cur["rows"].append(row)
    blocks = [b for b in blocks if re.search(re.escape(kre), b["name"])]
    b = blocks[min(which, len(blocks) - 1)]
    hdr, data = b["rows"][0], b["rows"][1:]
    ci = {h: i for i, h in enumerate(hdr)}
    addrs = [r[ci["Address"]] for r in data]
    if len(addrs) > 1 and addrs[0] in addrs[1:]:
        data = data[:addrs.index(addrs[0], 1)]
    sl = sass_lines(obj, re.escape(kre))

    def col(r, h):
        try:
            return float(r[ci[h]])
        except (ValueError, KeyError, IndexError):
            return 0.0
    tot_i = sum(col(r, "Instructions Executed") for r in data) or 1.0
    tot_s = sum(col(r, "# Samples") for r in data) or 1.0
    print(f"# {b['name']}: {len(data)} instr, {tot_i:.0f} warp instr, {tot_s:.0f} samples")
    cum = 0.0
    for i, r in enumerate(data):
        ie, te, sm = col(r, "Instructions Executed"), col(r, "Thread Instructions Executed"), col(r, "# Samples")
        cum += ie
        f, l = (sl[i][1], sl[i][2]) if i < len(sl) else ("?", 0)
        print(f"{i:5d} {ie / tot_i * 100:6.3f}% cum {cum / tot_i * 100:6.2f}% thr {te / max(1.0, ie):5.1f} smp {sm / tot_s * 100:5.2f}% "
              f"{f[:18]}:{l:<5d} {r[ci['Source']][:70]}")


if __name__ == "__main__":
    main()
