#!/bin/sh
# 4 CTAs/SM on the cfg3 stand-in: stack entries x subtree budget (smem per CTA = subtree + stacks + 18.5 KB <= 56 KB)
run() { printf "%-24s tb=%-6s cfg3 : " $1 $2; B2RT_VERBOSE=1 B2RT_LIB=$1 python tools/profile_frame.py --frames 3 --spp 16 --subdivide 1 --width 1920 --height 1080 --treelet-bytes $2 2>&1 | grep -o "[0-9] CTAs/SM\|frame.*" | sort -u | tr '\n' ' '; echo; }
run cuda-raytracer_b200/libb2rt.so 0
run cuda-raytracer_b200/libb2rt.so 22016
run build/st15/libb2rt.so 0
run build/st15/libb2rt.so 23040
run build/st15/libb2rt.so 22528
run build/st12/libb2rt.so 0
run build/st12/libb2rt.so 26624
printf "st15 cfg2: "; B2RT_LIB=build/st15/libb2rt.so python tools/profile_frame.py --frames 3 --spp 32 | tail -1
printf "st12 cfg2: "; B2RT_LIB=build/st12/libb2rt.so python tools/profile_frame.py --frames 3 --spp 32 | tail -1
