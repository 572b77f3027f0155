#!/bin/sh
# Build an experimental variant of libb2rt.so with extra nvcc flags, for A/B runs on the GPU box:
#   tools/build_variant.sh occ4 "-DB2RT_OCC4=4"   ->  build/occ4/libb2rt.so   (select with B2RT_LIB=build/occ4/libb2rt.so)
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
NAME=$1; shift
mkdir -p "$ROOT/build/$NAME/csrc" "$ROOT/build/include"
cp "$ROOT/include/b2rt.h" "$ROOT/build/include/"
cp "$ROOT"/cuda-raytracer_b200/csrc/*.cu "$ROOT"/cuda-raytracer_b200/csrc/*.cuh "$ROOT"/cuda-raytracer_b200/csrc/*.cpp \
   "$ROOT"/cuda-raytracer_b200/csrc/*.h "$ROOT"/cuda-raytracer_b200/csrc/Makefile "$ROOT/build/$NAME/csrc/"
make -C "$ROOT/build/$NAME/csrc" -j8 EXTRA="$*" >/dev/null
grep -h "registers" "$ROOT/build/$NAME/csrc/traverse.o.log" | sort | uniq -c
echo "built $ROOT/build/$NAME/libb2rt.so"
