#!/bin/sh
# A/B sweep of build variants (tools/build_variant.sh) and subtree budgets on cfg2 (32 spp) and the cfg3 stand-in (16 spp).
run() { # name lib treelet extra-args [env...]
  name=$1; lib=$2; tb=$3; extra=$4; shift 4
  printf "%-14s treelet=%-6s %-34s %s : " "$name" "$tb" "$extra" "$*"
  env B2RT_LIB=$lib "$@" timeout 120 python tools/profile_frame.py --frames 3 --treelet-bytes $tb $extra 2>&1 | tail -1
}
B=cuda-raytracer_b200/libb2rt.so
run pb256st7 $B 0 "--spp 32"
for v in pb0 st0 pb0st0 pb64 st4; do run $v build/$v/libb2rt.so 0 "--spp 32"; done
