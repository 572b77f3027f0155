#!/bin/sh
# PLOC search radius (build variants -DB2RT_PLOC_RADIUS=r): render workloads, a 457 K-triangle scene, the 10 M soup, build time
for lib in "$@"; do
  printf "%-32s cfg2 : " "$lib"; B2RT_LIB=$lib timeout 120 python tools/profile_frame.py --frames 3 --spp 32 2>&1 | tail -1
  printf "%-32s cfg3 : " "$lib"; B2RT_LIB=$lib timeout 120 python tools/profile_frame.py --frames 3 --spp 16 --subdivide 1 --width 1920 --height 1080 2>&1 | tail -1
  printf "%-32s cfg4c: " "$lib"; B2RT_LIB=$lib timeout 120 python tools/profile_frame.py --frames 3 --spp 16 --subdivide 2 --width 1920 --height 1080 2>&1 | tail -1
done
