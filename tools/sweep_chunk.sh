#!/bin/sh
# chunk size (rays per work item of a level >= 1 subtree queue) sweep: soup incoherent / coherent, cfg2, cfg3 stand-in
for c in 512 1024 2048 4096 8192; do
  printf "chunk=%-5s soup incoherent : " $c; B2RT_CHUNK_RAYS=$c python tools/profile_soup.py --builder gpu --mode 1 --repeats 3 2>&1 | tail -1
  printf "chunk=%-5s soup coherent   : " $c; B2RT_CHUNK_RAYS=$c python tools/profile_soup.py --builder gpu --mode 0 --repeats 3 2>&1 | tail -1
  printf "chunk=%-5s cfg2/32spp : " $c; B2RT_CHUNK_RAYS=$c python tools/profile_frame.py --frames 3 --spp 32 2>&1 | tail -1
  printf "chunk=%-5s cfg3/16spp : " $c; B2RT_CHUNK_RAYS=$c python tools/profile_frame.py --frames 3 --spp 16 --subdivide 1 --width 1920 --height 1080 2>&1 | tail -1
done
