#!/bin/sh
# renderer with the host (binned SAH) vs the device (LBVH) builder; B2RT_LARGE_PRIM = large-primitive threshold of the device builder
for cfg in "host 0.125" "gpu 0" "gpu 0.125" "gpu 0.25" "gpu 0.0625"; do
  set -- $cfg
  printf "builder=%-5s large=%-7s cfg2/32spp : " $1 $2; B2RT_BUILDER=$1 B2RT_LARGE_PRIM=$2 python tools/profile_frame.py --frames 3 --spp 32 2>&1 | tail -1
  printf "builder=%-5s large=%-7s cfg3/16spp : " $1 $2; B2RT_BUILDER=$1 B2RT_LARGE_PRIM=$2 python tools/profile_frame.py --frames 3 --spp 16 --subdivide 1 --width 1920 --height 1080 2>&1 | tail -1
done
B2RT_BUILDER=gpu python -m pytest tests -m gpu -x -q -k "radiance or device_built or waves" 2>&1 | tail -2
