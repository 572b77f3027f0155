#!/bin/sh
# quick A/B of library variants: one cfg2 (32 spp) and one cfg3 stand-in (16 spp) run per library
for lib in "$@"; do
  printf "%-36s cfg2 : " "$lib"
  B2RT_LIB=$lib timeout 120 python tools/profile_frame.py --frames 3 --spp 32 2>&1 | tail -1
  printf "%-36s cfg3 : " "$lib"
  B2RT_LIB=$lib timeout 120 python tools/profile_frame.py --frames 3 --spp 16 --subdivide 1 --width 1920 --height 1080 2>&1 | tail -1
done
