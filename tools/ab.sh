#!/bin/sh
# A/B of library variants on the GPU box: tools/ab.sh libA libB ...  (cfg2 at 32 spp, cfg3 stand-in at 16 spp)
for lib in "$@"; do
  for rep in 1 2; do
    printf "%-40s cfg2/32spp : " "$lib"
    B2RT_LIB=$lib timeout 120 python tools/profile_frame.py --frames 3 --spp 32 2>&1 | tail -1
  done
  printf "%-40s cfg3/16spp : " "$lib"
  B2RT_LIB=$lib timeout 120 python tools/profile_frame.py --frames 3 --spp 16 --subdivide 1 --width 1920 --height 1080 2>&1 | tail -1
done
