python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest17.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_pytest17.log
python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/bench_r02c_cfg3_n1.json 2> gpurun_out/bench_r02c_cfg3_n1.err; echo "bench cfg3 rc=$?"
python bench.py --workload cfg2 --steps 3 --warmup 3 --no-cpu > gpurun_out/bench_r02c_cfg2_n1.json 2> gpurun_out/bench_r02c_cfg2_n1.err; echo "bench cfg2 rc=$?"
python - <<'PY'
import json
for f in ("bench_r02c_cfg3_n1", "bench_r02c_cfg2_n1"):
    j = json.load(open(f"gpurun_out/{f}.json")); r = j["roofline"]
    print(f, "value %.0f" % j["value"], "ms/step %.2f" % j["ms_per_step"], "e2e %.0f" % j["e2e"]["value"], "e2e s/frame %.4f" % j["e2e"]["s_per_frame"],
          "roofline %.3f" % r["frac"], "share %.3f" % r["kernel_share_of_step"], "fp32 %.3f" % r["fp32"]["frac"])
    for k in ("level0", "deeper"):
        b = r["by_level"][k]; print("   ", k, "ms %.2f hbm_frac %.3f fp32_frac %.3f visits %.0fM" % (b["ms_per_frame"], b["hbm_frac"], b["fp32_frac"], b["visits"] / 1e6))
PY
