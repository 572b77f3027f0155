#!/bin/sh
# A/B recipe of a recorded experiment: shadow list without null rays (S = 1) vs the commit before it; build/head = that commit built with
# tools/build_variant.sh head "" from a checkout of it.  Run on the GPU box.
python -m pytest tests -m gpu -x -q -k "radiance or waves or full_size or sharding or median or renderer_on or cfg4 or dragon or slices_dense or filter" 2>&1 | tail -2
for lib in build/head/libb2rt.so cuda-raytracer_b200/libb2rt.so; do
  for rep in 1 2; do printf "%-34s cfg2/64spp overlap : " $lib; B2RT_LIB=$lib python tools/profile_frame.py --frames 3 --spp 64 --overlap 2>&1 | tail -1; done
  printf "%-34s cfg2/32spp serial  : " $lib; B2RT_LIB=$lib python tools/profile_frame.py --frames 3 --spp 32 2>&1 | tail -1
  printf "%-34s cfg3/32spp overlap : " $lib; B2RT_LIB=$lib python tools/profile_frame.py --frames 2 --spp 32 --overlap --subdivide 1 --width 1920 --height 1080 2>&1 | tail -1
done
