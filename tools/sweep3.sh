#!/bin/sh
# A/B sweep of build variants (tools/build_variant.sh) on cfg2 at 32 spp.
run() { # name lib treelet [env...]
  name=$1; lib=$2; tb=$3; shift 3
  printf "%-28s treelet=%-6s %s : " "$name" "$tb" "$*"
  env B2RT_LIB=$lib "$@" timeout 120 python tools/profile_frame.py --spp 32 --frames 3 --treelet-bytes $tb 2>&1 | tail -1
}
B=cuda-raytracer_b200/libb2rt.so
run base $B 24576
run base $B 16384
