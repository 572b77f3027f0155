#!/usr/bin/env python3
"""Markdown summary of an `ncu --set full` report: one row per captured launch with the counters north_star asks for
(achieved HBM GB/s, L2 / shared hit behaviour, warp execution efficiency, FP32 pipe utilisation, issue utilisation,
occupancy, registers).

  python tools/ncu_summary.py gpurun_out/prof.ncu-rep [kernel-substring] > profiles/xyz.md
"""
import csv
import io
import subprocess
import sys

METRICS = [
    ("gpu__time_duration.sum", "time us", "time"),
    ("launch__grid_size", "grid", 1.0),
    ("launch__registers_per_thread", "regs", 1.0),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy % (achieved)", 1.0),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %", 1.0),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "threads / warp instr (of 32)", 1.0),
    ("smsp__inst_executed.sum", "warp instr (M)", 1e-6),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "FMA pipe %", 1.0),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU pipe %", 1.0),
    ("dram__bytes_read.sum", "DRAM read MB", None),
    ("dram__bytes_write.sum", "DRAM write MB", None),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM % of peak", 1.0),
    ("lts__t_sector_hit_rate.pct", "L2 hit %", 1.0),
    ("l1tex__t_sector_hit_rate.pct", "L1 hit %", 1.0),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem wavefronts (M)", 1e-6),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts (M)", 1e-6),
]
UNIT_SCALE = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}


def main():
    rep = sys.argv[1]
    filt = sys.argv[2] if len(sys.argv) > 2 else ""
    raw = subprocess.check_output(["ncu", "-i", rep, "--page", "raw", "--csv"], text=True, stderr=subprocess.DEVNULL)
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ci = {h: i for i, h in enumerate(hdr)}
    data = [r for r in data if filt in r[ci["Kernel Name"]]]
    print(f"ncu --set full capture `{rep.split('/')[-1]}`, {len(data)} launch(es)" + (f" matching `{filt}`" if filt else "") + "\n")
    names = []
    for r in data:
        n = r[ci["Kernel Name"]].replace("b2rt::", "").replace("<unnamed>::", "").replace("void ", "").split("(")[0]
        names.append(n)
    print("| metric | " + " | ".join(f"#{i} `{n}`" for i, n in enumerate(names)) + " |")
    print("|---|" + "---|" * len(names))
    for key, label, scale in METRICS:
        if key not in ci:
            continue
        vals = []
        for r in data:
            try:
                v = float(r[ci[key]].replace(",", ""))
            except ValueError:
                vals.append("-")
                continue
            if scale == "time":
                sc = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(units[ci[key]], 1.0)
            else:
                sc = UNIT_SCALE.get(units[ci[key]], 1.0) if scale is None else scale
            v *= sc
            vals.append(f"{v:.2f}" if abs(v) < 100 else f"{v:.0f}")
        print(f"| {label} | " + " | ".join(vals) + " |")
    # achieved DRAM GB/s
    if "dram__bytes_read.sum" in ci:
        vals = []
        for r in data:
            rb = float(r[ci["dram__bytes_read.sum"]].replace(",", "")) * UNIT_SCALE.get(units[ci["dram__bytes_read.sum"]], 1.0)
            wb = float(r[ci["dram__bytes_write.sum"]].replace(",", "")) * UNIT_SCALE.get(units[ci["dram__bytes_write.sum"]], 1.0)
            t = float(r[ci["gpu__time_duration.sum"]].replace(",", ""))
            tu = units[ci["gpu__time_duration.sum"]]
            t_us = t * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(tu, 1.0)
            vals.append(f"{(rb + wb) / t_us * 1e3:.0f}")   # MB/us = TB/s -> GB/s
        print("| achieved DRAM GB/s | " + " | ".join(vals) + " |")


if __name__ == "__main__":
    main()
