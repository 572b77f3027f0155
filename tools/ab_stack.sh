#!/bin/sh
# A/B recipe of a recorded experiment (per-ray stack of 16 vs 24 entries).  Needs: tools/build_variant.sh st24 "-DB2RT_STACK4=24".  Run on the GPU box.
for lib in cuda-raytracer_b200/libb2rt.so build/st24/libb2rt.so; do
 for b in host gpu; do
  printf "%-32s builder=%-4s cfg2/32spp : " $lib $b; B2RT_LIB=$lib B2RT_BUILDER=$b python tools/profile_frame.py --frames 3 --spp 32 2>&1 | tail -1
  printf "%-32s builder=%-4s cfg3/16spp : " $lib $b; B2RT_LIB=$lib B2RT_BUILDER=$b python tools/profile_frame.py --frames 3 --spp 16 --subdivide 1 --width 1920 --height 1080 2>&1 | tail -1
 done
done
