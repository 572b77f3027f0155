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

# ---- second reported baseline: the reference's own CUDA renderer, headless (oracle/ref_cuda_driver.cpp replaces
# src/cudaMain.cpp + src/display.cpp).  src/cudaRenderer.cu and the host sources its loadScene() needs are compiled
# UNMODIFIED from where they lie; objects go to a scratch directory, the binary and the two scene files it is run on
# go to oracle/_ref/ (git-ignored, travels to the GPU box).  oracle/ref_glu_protos.h supplies the <GL/glu.h>
# prototypes the reference's draw code wants (never called).  Skipped without nvcc.
if command -v nvcc >/dev/null 2>&1 || [ -x /usr/local/cuda/bin/nvcc ]; then
  NVCC=$(command -v nvcc || echo /usr/local/cuda/bin/nvcc)
  TMP=$(mktemp -d)
  INC="-I $REF/CMU462/include -I $REF/CMU462/include/CMU462 -I $REF/src -I /usr/local/cuda/include"
  for f in "$REF"/src/bvh.cpp "$REF"/src/bbox.cpp "$REF"/src/bsdf.cpp "$REF"/src/camera.cpp "$REF"/src/sampler.cpp \
           "$REF"/src/halfEdgeMesh.cpp "$REF"/src/meshEdit.cpp "$REF"/src/static_scene/*.cpp "$REF"/src/collada/*.cpp \
           "$REF"/src/dynamic_scene/*.cpp "$REF"/CMU462/src/vector3D.cpp "$REF"/CMU462/src/vector2D.cpp \
           "$REF"/CMU462/src/vector4D.cpp "$REF"/CMU462/src/matrix3x3.cpp "$REF"/CMU462/src/matrix4x4.cpp \
           "$REF"/CMU462/src/color.cpp "$REF"/CMU462/src/spectrum.cpp "$REF"/CMU462/src/tinyxml2.cpp \
           "$REF"/CMU462/src/lodepng.cpp "$REF"/CMU462/src/quaternion.cpp "$REF"/CMU462/src/complex.cpp; do
    case "$f" in */static_scene/bvh.cpp) continue ;; esac      # stale duplicate of src/bvh.cpp, not in the reference's build
    g++ -O2 -std=gnu++11 -include "$HERE/ref_glu_protos.h" -DGLEW_NO_GLU -w $INC -c "$f" -o "$TMP/$(echo "$f" | tr '/' '_').o" &
  done
  g++ -O2 -std=gnu++11 -include "$HERE/ref_glu_protos.h" -DGLEW_NO_GLU -w $INC -c "$HERE/ref_cuda_driver.cpp" -o "$TMP/driver.o" &
  "$NVCC" -std=c++14 -gencode arch=compute_100a,code=sm_100a -O3 --pre-include "$HERE/ref_glu_protos.h" -DGLEW_NO_GLU -w $INC \
      -c "$REF/src/cudaRenderer.cu" -o "$TMP/cudaRenderer.o"
  wait
  "$NVCC" -gencode arch=compute_100a,code=sm_100a -o "$HERE/_ref/ref_cuda_render" "$TMP"/*.o -Xlinker --unresolved-symbols=ignore-all -lcurand
  mkdir -p "$HERE/_ref/media"
  install -m 644 "$REF/media/pathtracer/advanced/CBbunny.dae" "$REF/media/pathtracer/advanced/CBcoil.dae" "$HERE/_ref/media/"
  rm -rf "$TMP"
  echo "built $HERE/_ref/ref_cuda_render"
fi
