#!/bin/sh
# Multi-GPU measurements beyond the driver's cfg2 weak-scaling run (one box, `gpurun --gpus 8`):
#   BASELINE configs[2] stand-in, 256 spp sharded over 8 / 4 / 2 GPUs (strong scaling; N=1 is run on a 1-GPU box),
#   BASELINE configs[3] stand-in, 1024 spp over 8 GPUs with the NCCL accumulation reduce, and cfg2 at 8 GPUs.
run() { # N workload steps warmup
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $((29500 + $1)) \
    bench.py --gpus $1 --workload $2 --steps $3 --warmup $4 --no-cpu > gpurun_out/bench_$2_n$1.json 2> gpurun_out/bench_$2_n$1.err
  echo "rc=$? $2 N=$1: $(head -c 420 gpurun_out/bench_$2_n$1.json)"
}
mkdir -p gpurun_out
run 8 cfg3 2 1
run 8 cfg4 1 1
run 4 cfg3 2 1
run 2 cfg3 2 1
run 8 cfg2 3 3
