#!/bin/sh
# parameter sweep on the cfg2 frame at 16 spp (frame ms / Mrays/s): BVH width x subtree budget x wave size
for w in 4 8; do for tb in 16384 32768 49152 65536 98304 163840; do
  printf "W=%s treelet=%s : " $w $tb; timeout 60 python tools/profile_frame.py --spp 16 --frames 3 --bvh-width $w --treelet-bytes $tb | tail -1
done; done
for wave in 1000000 2000000 8000000; do printf "wave=%s : " $wave; timeout 60 python tools/profile_frame.py --spp 16 --frames 3 --wave $wave | tail -1; done
