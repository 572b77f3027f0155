#!/bin/sh
run() { # name lib treelet extra-args [env...]
  name=$1; lib=$2; tb=$3; extra=$4; shift 4
  printf "%-14s treelet=%-6s %-50s %s : " "$name" "$tb" "$extra" "$*"
  env B2RT_LIB=$lib "$@" timeout 120 python tools/profile_frame.py --frames 3 --treelet-bytes $tb $extra 2>&1 | tail -1
}
B=cuda-raytracer_b200/libb2rt.so
run count $B 0 "--spp 32"
run count $B 0 "--spp 16 --subdivide 1 --width 1920 --height 1080"
