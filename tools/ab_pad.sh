#!/bin/sh
# A/B recipe of a recorded experiment (node stride 128 vs 144 B).  Needs the variant first: tools/build_variant.sh pad16 "-DB2RT_NODE_PAD=16"
# (144 B is the default now; build the other side with -DB2RT_NODE_PAD=0).  Run on the GPU box.
B2RT_LIB=build/pad16/libb2rt.so python -m pytest tests -m gpu -x -q 2>&1 | tail -2
for lib in cuda-raytracer_b200/libb2rt.so build/pad16/libb2rt.so; do
  printf "%-36s soup incoherent : " $lib; B2RT_LIB=$lib python tools/profile_soup.py --builder gpu --mode 1 --repeats 3 2>&1 | tail -1
  printf "%-36s soup coherent   : " $lib; B2RT_LIB=$lib python tools/profile_soup.py --builder gpu --mode 0 --repeats 3 2>&1 | tail -1
done
tools/ab.sh cuda-raytracer_b200/libb2rt.so build/pad16/libb2rt.so
