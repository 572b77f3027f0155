# final checks of the round: smoke(), full GPU suite, both bench arms, launch lists
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest19.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_pytest19.log
python bench.py --steps 3 --warmup 3 > gpurun_out/bench_r02e_cfg3_n1.json 2> gpurun_out/bench_r02e_cfg3_n1.err; echo "bench cfg3 rc=$?"
python bench.py --workload cfg2 --steps 3 --warmup 3 > gpurun_out/bench_r02e_cfg2_n1.json 2> gpurun_out/bench_r02e_cfg2_n1.err; echo "bench cfg2 rc=$?"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r02e_ref.json 2> gpurun_out/bench_r02e_ref.err; echo "bench ref rc=$?"; cat gpurun_out/bench_r02e_ref.json | cut -c1-400
python bench.py --steps 1 --warmup 0 --no-cpu --spp 32 > gpurun_out/bench_spp32.json 2>/dev/null && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 1000 --csv --log-file gpurun_out/launches_r02_bench_cfg3.csv python bench.py --steps 1 --warmup 0 --no-cpu --spp 32 > gpurun_out/ncu_bench_cfg3.log 2>&1
python tools/launch_summary.py gpurun_out/launches_r02_bench_cfg3.csv > gpurun_out/r02_launch_summary_cfg3.txt; head -14 gpurun_out/r02_launch_summary_cfg3.txt
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 1000 --csv --log-file gpurun_out/launches_r02_cfg2_64spp.csv python tools/profile_frame.py --frames 1 --spp 64 > gpurun_out/ncu_cfg2.log 2>&1
python tools/launch_summary.py gpurun_out/launches_r02_cfg2_64spp.csv > gpurun_out/r02_launch_summary_cfg2.txt; head -12 gpurun_out/r02_launch_summary_cfg2.txt
python - <<'PY'
import json
for f in ("bench_r02e_cfg3_n1", "bench_r02e_cfg2_n1"):
    j = json.load(open(f"gpurun_out/{f}.json")); r = j["roofline"]
    print(f, "value %.0f" % j["value"], "ms/step %.2f" % j["ms_per_step"], "e2e %.0f" % j["e2e"]["value"], "e2e s/frame %.4f" % j["e2e"]["s_per_frame"],
          "roofline %.3f" % r["frac"], "share %.3f" % r["kernel_share_of_step"], "fp32 %.3f" % r["fp32"]["frac"], "cpu", (j.get("cpu_baseline") or {}), "clocks", j["clocks"])
    for k in ("level0", "deeper"):
        b = r["by_level"][k]; print("   ", k, "ms %.2f hbm_frac %.3f fp32_frac %.3f visits %.0fM" % (b["ms_per_frame"], b["hbm_frac"], b["fp32_frac"], b["visits"] / 1e6))
PY
