python -m pytest tests -m gpu -x -q 2>&1 | tail -n 8
python bench.py --steps 3 --warmup 3 > gpurun_out/bench_r02_cfg3_n1.json 2> gpurun_out/bench_r02_cfg3_n1.err; echo "bench cfg3 rc=$?"; tail -c 600 gpurun_out/bench_r02_cfg3_n1.err
python bench.py --workload cfg2 --steps 3 --warmup 3 > gpurun_out/bench_r02_cfg2_n1.json 2> gpurun_out/bench_r02_cfg2_n1.err; echo "bench cfg2 rc=$?"
python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/bench_r02_ref.json 2> gpurun_out/bench_r02_ref.err; echo "ref rc=$?"
python - <<'PY'
import json
for f in ("bench_r02_cfg3_n1", "bench_r02_cfg2_n1", "bench_r02_ref"):
    try:
        j = json.load(open(f"gpurun_out/{f}.json"))
    except Exception as e:
        print(f, "unreadable", e); continue
    r = j.get("roofline") or {}
    print(f, "value %.0f" % j["value"], "ms/step %.2f" % j["ms_per_step"], "e2e %.0f" % j["e2e"]["value"], "e2e s/frame", j["e2e"].get("s_per_frame"),
          "roofline %.3f" % r.get("frac", 0), "share %.3f" % r.get("kernel_share_of_step", 0), "fp32", (r.get("fp32") or {}).get("peak_tflops"), (r.get("fp32") or {}).get("frac"),
          "cpu", (j.get("cpu_baseline") or {}).get("value"))
    if r.get("by_level"):
        for k in ("level0", "deeper"):
            b = r["by_level"][k]; print("   ", k, "ms %.2f hbm_frac %.3f fp32_frac %.3f visits %.0fM" % (b["ms_per_frame"], b["hbm_frac"], b["fp32_frac"], b["visits"] / 1e6))
PY
