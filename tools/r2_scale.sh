#!/bin/sh
# strong-scaling run of the default workload (cfg3 stand-in, 256 spp over N ranks) on N GPUs of one box
N=$1
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 3 --warmup 3 --no-cpu > gpurun_out/bench_r02f_cfg3_n$N.json 2> gpurun_out/bench_r02f_cfg3_n$N.err
echo "rc=$?"; tail -c 600 gpurun_out/bench_r02f_cfg3_n$N.err
python - <<PY
import json
j = json.load(open("gpurun_out/bench_r02f_cfg3_n$N.json"))
print("N", j["n_gpus"], "value %.0f" % j["value"], "ms/step %.2f" % j["ms_per_step"], "e2e %.0f" % j["e2e"]["value"], "e2e s/frame %.4f" % j["e2e"]["s_per_frame"], j.get("collective"), j["clocks"])
PY
