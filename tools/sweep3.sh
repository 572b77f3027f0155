#!/bin/sh
# A/B sweep of traversal-kernel build variants (tools/build_variant.sh) x subtree budgets on cfg2 at 32 spp.
run() { # name lib treelet [env...]
  name=$1; lib=$2; tb=$3; shift 3
  printf "%-28s treelet=%-6s %s : " "$name" "$tb" "$*"
  env B2RT_LIB=$lib "$@" timeout 120 python tools/profile_frame.py --spp 32 --frames 3 --treelet-bytes $tb 2>&1 | tail -1
}
B=cuda-raytracer_b200/libb2rt.so
run base $B 40960
run base $B 32768
run mi4 build/mi4/libb2rt.so 32768
run mi16 build/mi16/libb2rt.so 32768
run occ4s16 build/occ4s16/libb2rt.so 24576
run occ4s16 build/occ4s16/libb2rt.so 20480
run occ4s16 build/occ4s16/libb2rt.so 16384
run occ4s16mi4 build/occ4s16mi4/libb2rt.so 24576
run occ4s16mi12 build/occ4s16mi12/libb2rt.so 24576
