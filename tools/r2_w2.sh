#!/bin/sh
# W = 2 against W = 4 on the render workloads (host builder; cfg2 32 spp, cfg3 stand-in 16 spp), leaf sizes 2..4
for w in 4 2; do for leaf in 0 2 4; do
  printf "W=%s leaf=%s cfg2 : " $w $leaf
  B2RT_BUILDER=host timeout 120 python tools/profile_frame.py --frames 3 --spp 32 --bvh-width $w --max-leaf $leaf 2>&1 | tail -1
  printf "W=%s leaf=%s cfg3 : " $w $leaf
  B2RT_BUILDER=host timeout 120 python tools/profile_frame.py --frames 3 --spp 16 --subdivide 1 --width 1920 --height 1080 --bvh-width $w --max-leaf $leaf 2>&1 | tail -1
done; done
