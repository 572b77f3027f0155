#!/bin/sh
# A/B sweep of traversal-kernel build variants (tools/build_variant.sh) x subtree budgets on cfg2 at 32 spp.
run() { # name lib treelet [env...]
  name=$1; lib=$2; tb=$3; shift 3
  printf "%-28s treelet=%-6s %s : " "$name" "$tb" "$*"
  env B2RT_LIB=$lib "$@" timeout 120 python tools/profile_frame.py --spp 32 --frames 3 --treelet-bytes $tb 2>&1 | tail -1
}
B=cuda-raytracer_b200/libb2rt.so
B2RT_VERBOSE=1 python tools/profile_frame.py --spp 8 2>&1 | grep "b2rt:" | head -2
run base_mi12 $B 24576
run base_mi12 $B 20480
run mi1 build/mi1/libb2rt.so 24576
run mi4 build/mi4/libb2rt.so 24576
run mi8 build/mi8/libb2rt.so 24576
run mi16 build/mi16/libb2rt.so 24576
