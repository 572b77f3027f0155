// b2rt_load_dae: COLLADA loader ("next" row).  Placeholder until the C++ loader lands:
// scenes are converted offline by tools/dae2scene.py and loaded with b2rt_scene_load.
#include "b2rt_internal.h"

extern "C" int b2rt_load_dae(const char* path, b2rt_scene_file** out) {
  (void)path;
  if (out) *out = nullptr;
  b2rt::set_error("b2rt_load_dae: not built into this library yet; convert with tools/dae2scene.py and use b2rt_scene_load");
  return B2RT_ERR_IO;
}
