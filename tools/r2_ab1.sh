#!/bin/sh
# round 2, first GPU call: parity suite on the decoupled-leaf traversal kernel, then A/B against the round-1 build
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest1.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_pytest1.log
{
sh tools/ab.sh build/r1/libb2rt.so cuda-raytracer_b200/libb2rt.so build/idle8/libb2rt.so build/idle12/libb2rt.so build/idle16/libb2rt.so build/idle26/libb2rt.so build/take2/libb2rt.so
for tb in 16384 20480 24576; do
  for lib in build/r1/libb2rt.so cuda-raytracer_b200/libb2rt.so build/idle12/libb2rt.so; do
    printf "%-40s cfg3/16spp tb=%s : " "$lib" "$tb"
    B2RT_LIB=$lib timeout 120 python tools/profile_frame.py --frames 3 --spp 16 --subdivide 1 --width 1920 --height 1080 --treelet-bytes $tb 2>&1 | tail -1
  done
done
} > gpurun_out/r2_ab1.txt 2>&1
cat gpurun_out/r2_ab1.txt
