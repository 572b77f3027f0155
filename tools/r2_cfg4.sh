#!/bin/sh
# BASELINE configs[3] stand-in (457 K glass triangles + mirror spheres, 3840x2160, 1024 spp over N ranks, strong scaling)
N=$1
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 bench.py --workload cfg4 --gpus $N --steps 2 --warmup 1 --no-cpu > gpurun_out/bench_r02_cfg4_n$N.json 2> gpurun_out/bench_r02_cfg4_n$N.err
echo "rc=$?"; tail -c 300 gpurun_out/bench_r02_cfg4_n$N.err
python - <<PY
import json
j = json.load(open("gpurun_out/bench_r02_cfg4_n$N.json"))
print("N", j["n_gpus"], "value %.0f" % j["value"], "ms/step %.2f" % j["ms_per_step"], "e2e %.0f" % j["e2e"]["value"], "e2e s/frame %.4f" % j["e2e"]["s_per_frame"], j["clocks"], j["bvh"])
PY
