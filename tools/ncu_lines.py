#!/usr/bin/env python3
"""Per-source-line attribution of an `ncu --set full --import-source on` capture.

`ncu --page source --csv` only lists SASS; this joins it with `nvdisasm -g` line info of the object file the
kernel came from, by instruction order inside the kernel, and prints stall samples / executed instructions / average
active threads per CUDA source line plus the kernel-wide stall-reason mix.

  python tools/ncu_lines.py gpurun_out/prof.ncu-rep cuda-raytracer_b200/csrc/traverse.o 'k_traverse<(int)4, (bool)0' [launch#]
"""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile


def sass_lines(obj, kernel_re):
    """[(sass text, file, line)] for the first function of `obj` whose demangled name matches kernel_re."""
    tmp = tempfile.mkdtemp()
    subprocess.check_call(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, stdout=subprocess.DEVNULL)
    cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
    txt = subprocess.check_output(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], text=True, stderr=subprocess.DEVNULL)
    out, cur, take = [], None, False
    for ln in txt.splitlines():
        m = re.match(r"\s*\.text\.(\S+):", ln)
        if m:
            if take and out:
                break
            name = subprocess.check_output(["cu++filt", m.group(1)], text=True).strip()
            take = re.search(kernel_re, name) is not None
            cur = None
            continue
        if not take:
            continue
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
        if m:
            out.append((m.group(2).strip(), cur[0] if cur else "?", cur[1] if cur else 0))
    return out


def main():
    rep, obj, kre = sys.argv[1], sys.argv[2], sys.argv[3]
    which = int(sys.argv[4]) if len(sys.argv) > 4 else 0
    kre_plain = re.escape(kre)
    raw = subprocess.check_output(["ncu", "-i", rep, "--page", "source", "--csv"], text=True, stderr=subprocess.DEVNULL)
    # the CSV is a sequence of per-kernel blocks, each starting with a "Kernel Name" row
    blocks, cur = [], None
    for row in csv.reader(io.StringIO(raw)):
        if row and row[0] == "Kernel Name":
            cur = {"name": row[1], "rows": []}
            blocks.append(cur)
        elif cur is not None and row:
            cur["rows"].append(row)
    blocks = [b for b in blocks if re.search(kre_plain, b["name"])]
    if not blocks:
        sys.exit("no kernel matching " + kre)
    b = blocks[min(which, len(blocks) - 1)]
    hdr, data = b["rows"][0], b["rows"][1:]
    ci = {h: i for i, h in enumerate(hdr)}
    # ncu lists the instructions twice when the report holds two passes; keep the first copy
    addrs = [r[ci["Address"]] for r in data]
    if len(addrs) > 1 and addrs[0] in addrs[1:]:
        data = data[:addrs.index(addrs[0], 1)]
    sl = sass_lines(obj, kre_plain)
    print(f"# {b['name']}\n# {len(data)} SASS instructions in the report, {len(sl)} in {obj}")
    if len(data) != len(sl):
        print("# WARNING: instruction counts differ; the object file is not the one that was profiled")

    def col(r, h):
        try:
            return float(r[ci[h]])
        except (ValueError, KeyError, IndexError):
            return 0.0
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    tot_s = sum(col(r, "# Samples") for r in data) or 1.0
    tot_i = sum(col(r, "Instructions Executed") for r in data) or 1.0
    tot_t = sum(col(r, "Thread Instructions Executed") for r in data)
    print(f"# samples {tot_s:.0f}, warp instructions {tot_i:.0f}, avg active threads {tot_t / tot_i:.1f}")
    agg = sorted(((sum(col(r, s) for r in data), s) for s in stalls), reverse=True)
    print("# stall mix: " + ", ".join(f"{s[6:]} {v / tot_s * 100:.1f}%" for v, s in agg[:9]))
    by = collections.OrderedDict()
    for i, r in enumerate(data):
        key = (sl[i][1], sl[i][2]) if i < len(sl) else ("?", 0)
        e = by.setdefault(key, [0.0, 0.0, 0.0, collections.Counter(), 0.0, 0.0])
        e[0] += col(r, "# Samples"); e[1] += col(r, "Instructions Executed"); e[2] += col(r, "Thread Instructions Executed")
        for s in stalls:
            e[3][s] += col(r, s)
        e[4] += col(r, "L1 Wavefronts Shared"); e[5] += col(r, "L2 Theoretical Sectors Local")
    src = {}
    print(f"{'file:line':>22s} {'smp%':>6s} {'inst%':>6s} {'thr':>5s} {'smemWF':>9s} {'locL2':>8s}  top stall | source")
    for (f, l), e in sorted(by.items(), key=lambda kv: (kv[0][0], kv[0][1])):
        if e[0] / tot_s < 0.003 and e[1] / tot_i < 0.004:
            continue
        if f not in src:
            p = os.path.join(os.path.dirname(obj), f)
            src[f] = open(p).read().splitlines() if os.path.exists(p) else []
        text = src[f][l - 1].strip()[:90] if 0 < l <= len(src[f]) else ""
        top = e[3].most_common(1)[0][0][6:] if e[0] else "-"
        print(f"{f[:16] + ':' + str(l):>22s} {e[0] / tot_s * 100:6.1f} {e[1] / tot_i * 100:6.1f} {e[2] / max(1.0, e[1]):5.1f} "
              f"{e[4]:9.0f} {e[5]:8.0f}  {top:<18s}| {text}")


if __name__ == "__main__":
    main()
