// Headless render with the Scotty3D call sequence (Application::render_scene + set_up_pathtracer,
// src/application.cpp:1979-1989, 1593-1603; CLI flags -s -l -m -w of src/main.cpp:78-105) on top of the C ABI.
//   render_scene [-s ns_aa] [-l ns_area_light] [-m max_ray_depth] [-r WxH] [-w out.png] scene.{b2s,dae}
#include <cstdlib>
#include <iostream>

#include "../shim/scotty_shim.h"

int main(int argc, char** argv) {
  size_t ns_aa = 16, ns_area = 1, depth = 4;
  uint32_t w = 640, h = 480;
  std::string out = "out.png", scene;
  for (int i = 1; i < argc; ++i) {
    std::string a = argv[i];
    if (a == "-s" && i + 1 < argc) ns_aa = atoi(argv[++i]);
    else if (a == "-l" && i + 1 < argc) ns_area = atoi(argv[++i]);
    else if (a == "-m" && i + 1 < argc) depth = atoi(argv[++i]);
    else if (a == "-w" && i + 1 < argc) out = argv[++i];
    else if (a == "-r" && i + 1 < argc) { if (sscanf(argv[++i], "%ux%u", &w, &h) != 2) return 2; }
    else scene = a;
  }
  if (scene.empty()) { std::cerr << "usage: render_scene [-s spp] [-l light samples] [-m depth] [-r WxH] [-w out.png] scene.b2s\n"; return 2; }
  try {
    b2rt_shim::SceneFile sf(scene);
    b2rt_shim::PathTracer pt(ns_aa, depth, ns_area);
    b2rt_camera cam = sf.camera(w, h);
    pt.set_camera(&cam);
    pt.set_scene(sf.desc());
    pt.set_frame_size(w, h);
    pt.start_raytracing();
    while (!pt.is_done()) std::this_thread::sleep_for(std::chrono::milliseconds(5));   // application.cpp:1984
    pt.save_image(out);
    b2rt_stats st = pt.stats();
    const double rays = (double)(st.rays_camera + st.rays_bounce + st.rays_shadow);
    std::cout << "rendered " << scene << " " << w << "x" << h << " " << ns_aa << " spp depth " << depth << ": " << st.ms_total
              << " ms, " << rays / st.ms_total / 1e3 << " Mrays/s, BVH " << st.bvh_nodes << " nodes / " << st.bvh_subtrees
              << " subtrees / " << st.bvh_levels << " levels -> " << out << "\n";
  } catch (const std::exception& e) {
    std::cerr << e.what() << "\n";
    return 1;
  }
  return 0;
}
