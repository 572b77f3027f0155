#!/bin/sh
# leaf-size sweep on cfg2 (32 spp) and the cfg3 stand-in (16 spp)
run() { # name lib treelet extra-args [env...]
  name=$1; lib=$2; tb=$3; extra=$4; shift 4
  printf "%-14s treelet=%-6s %-50s %s : " "$name" "$tb" "$extra" "$*"
  env B2RT_LIB=$lib "$@" timeout 120 python tools/profile_frame.py --frames 3 --treelet-bytes $tb $extra 2>&1 | tail -1
}
B=cuda-raytracer_b200/libb2rt.so
for l in 1 2 3 4 6; do run leaf$l $B 0 "--spp 32 --max-leaf $l"; done
for l in 2 3 4; do run leaf$l $B 0 "--spp 16 --subdivide 1 --width 1920 --height 1080 --max-leaf $l"; done
