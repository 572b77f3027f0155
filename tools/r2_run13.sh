python bench.py --steps 3 --warmup 3 > gpurun_out/bench_r02b_cfg3_n1.json 2> gpurun_out/bench_r02b_cfg3_n1.err; echo "bench cfg3 rc=$?"
python bench.py --workload cfg2 --steps 3 --warmup 3 > gpurun_out/bench_r02b_cfg2_n1.json 2> gpurun_out/bench_r02b_cfg2_n1.err; echo "bench cfg2 rc=$?"
python bench.py --steps 1 --warmup 0 --no-cpu --spp 32 > gpurun_out/bench_spp32.json 2>/dev/null && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 1000 --csv --log-file gpurun_out/launches_r02_bench_cfg3.csv python bench.py --steps 1 --warmup 0 --no-cpu --spp 32 > gpurun_out/ncu_bench_cfg3.log 2>&1
python tools/launch_summary.py gpurun_out/launches_r02_bench_cfg3.csv
python - <<'PY'
import json
for f in ("bench_r02b_cfg3_n1", "bench_r02b_cfg2_n1"):
    j = json.load(open(f"gpurun_out/{f}.json")); r = j["roofline"]
    print(f, "value %.0f" % j["value"], "ms/step %.2f" % j["ms_per_step"], "e2e %.0f" % j["e2e"]["value"], "e2e s/frame %.4f" % j["e2e"]["s_per_frame"],
          "roofline %.3f" % r["frac"], "share %.3f" % r["kernel_share_of_step"], "fp32 %.3f" % r["fp32"]["frac"], "cpu", (j.get("cpu_baseline") or {}).get("value"), "bvh", j["bvh"])
    for k in ("level0", "deeper"):
        b = r["by_level"][k]; print("   ", k, "ms %.2f hbm_frac %.3f fp32_frac %.3f visits %.0fM" % (b["ms_per_frame"], b["hbm_frac"], b["fp32_frac"], b["visits"] / 1e6))
PY
