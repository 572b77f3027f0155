#!/bin/sh
# shadow trace of bounce b on a second stream next to the closest-hit trace of bounce b + 1 (B2RT_OVERLAP=1) vs one stream
B2RT_OVERLAP=1 python -m pytest tests -m gpu -x -q -k "radiance or waves or full_size or sharding or median or renderer_on" 2>&1 | tail -2
for ov in 0 1; do
  for rep in 1 2; do printf "overlap=%s cfg2/32spp : " $ov; B2RT_OVERLAP=$ov python tools/profile_frame.py --frames 3 --spp 32 2>&1 | tail -1; done
  printf "overlap=%s cfg2/64spp : " $ov; B2RT_OVERLAP=$ov python tools/profile_frame.py --frames 3 --spp 64 2>&1 | tail -1
  printf "overlap=%s cfg3/16spp : " $ov; B2RT_OVERLAP=$ov python tools/profile_frame.py --frames 3 --spp 16 --subdivide 1 --width 1920 --height 1080 2>&1 | tail -1
done
