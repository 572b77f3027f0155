#!/bin/sh
# ORACLE / TEST INFRASTRUCTURE.  Compiles the reference's OWN BVH builder (src/bvh.cpp, src/bbox.cpp and
# the CMU462 math sources it needs) from where they lie under $REF (default /root/reference) together
# with oracle/ref_bvh_driver.cpp into oracle/_ref/ref_bvh_dump.  No reference source is copied.
# -include cstdint : gcc 13 needs it for CMU462/spectrum.h; GLEW_NO_GLU : bundled glew.h wants GL/glu.h;
# --unresolved-symbols=ignore-all : bbox.cpp's BBox::draw references OpenGL entry points never called here.
set -e
REF=${REF:-/root/reference}
HERE=$(cd "$(dirname "$0")" && pwd)
[ -d "$REF/src" ] || { echo "reference not present at $REF; keeping prebuilt oracle/_ref"; exit 0; }
mkdir -p "$HERE/_ref"
g++ -O2 -std=gnu++11 -include cstdint -DGLEW_NO_GLU -w \
    -I "$REF/CMU462/include" -I "$REF/CMU462/include/CMU462" -I "$REF/src" \
    "$HERE/ref_bvh_driver.cpp" "$REF/src/bvh.cpp" "$REF/src/bbox.cpp" \
    "$REF/CMU462/src/vector3D.cpp" "$REF/CMU462/src/vector2D.cpp" "$REF/CMU462/src/vector4D.cpp" \
    "$REF/CMU462/src/matrix3x3.cpp" "$REF/CMU462/src/matrix4x4.cpp" "$REF/CMU462/src/color.cpp" "$REF/CMU462/src/spectrum.cpp" \
    -Wl,--unresolved-symbols=ignore-all -o "$HERE/_ref/ref_bvh_dump"
echo "built $HERE/_ref/ref_bvh_dump"
