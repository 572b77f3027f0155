// Headless render with the Scotty3D call sequence (Application::render_scene + set_up_pathtracer,
// src/application.cpp:1979-1989, 1593-1603; CLI flags -s -l -m -w of src/main.cpp:78-105) on top of the C ABI.
//   render_scene [-s ns_aa] [-l ns_area_light] [-m max_ray_depth] [-r WxH] [-w out.png] [-x out.exr] [-g gpus] scene.{b2s,dae}
// -g N (no reference equivalent: the reference is single-GPU): the job's ns_aa samples are dealt round-robin to N
// GPUs of this process (sample s goes to GPU s mod N), the scene is replicated, and the per-GPU accumulation buffers
// are combined on GPU 0 with ONE NCCL reduce (b2rt_comm_create_all + b2rt_reduce_accum_all).
#include <cstdlib>
#include <iostream>
#include <memory>

#include "../shim/scotty_shim.h"

int main(int argc, char** argv) {
  size_t ns_aa = 16, ns_area = 1, depth = 4;
  uint32_t w = 640, h = 480;
  int gpus = 1;
  std::string out = "out.png", exr, scene;
  for (int i = 1; i < argc; ++i) {
    std::string a = argv[i];
    if (a == "-s" && i + 1 < argc) ns_aa = atoi(argv[++i]);
    else if (a == "-l" && i + 1 < argc) ns_area = atoi(argv[++i]);
    else if (a == "-m" && i + 1 < argc) depth = atoi(argv[++i]);
    else if (a == "-w" && i + 1 < argc) out = argv[++i];
    else if (a == "-x" && i + 1 < argc) exr = argv[++i];
    else if (a == "-g" && i + 1 < argc) gpus = atoi(argv[++i]);
    else if (a == "-r" && i + 1 < argc) { if (sscanf(argv[++i], "%ux%u", &w, &h) != 2) return 2; }
    else scene = a;
  }
  if (scene.empty() || gpus < 1) {
    std::cerr << "usage: render_scene [-s spp] [-l light samples] [-m depth] [-r WxH] [-w out.png] [-x out.exr] [-g gpus] scene.{b2s,dae}\n";
    return 2;
  }
  try {
    b2rt_shim::SceneFile sf(scene);
    b2rt_camera cam = sf.camera(w, h);
    if (gpus > b2rt_device_count()) throw std::runtime_error("b2rt: fewer CUDA devices than -g asks for (there is no CPU fallback)");
    std::vector<std::unique_ptr<b2rt_shim::PathTracer>> pts;
    for (int g = 0; g < gpus; ++g) {
      b2rt_config cfg;
      memset(&cfg, 0, sizeof cfg);
      cfg.ns_aa = (uint32_t)(ns_aa / gpus + ((size_t)g < ns_aa % gpus ? 1 : 0));   // this GPU's share of the samples
      cfg.max_ray_depth = (uint32_t)depth; cfg.ns_area_light = (uint32_t)ns_area;
      cfg.device = g; cfg.sample_first = (uint32_t)g; cfg.sample_stride = (uint32_t)gpus;
      if (cfg.ns_aa == 0) throw std::runtime_error("b2rt: fewer samples than GPUs");
      pts.emplace_back(new b2rt_shim::PathTracer(cfg));
      pts.back()->set_camera(&cam);
      pts.back()->set_scene(sf.desc());
      pts.back()->set_frame_size(w, h);
    }
    for (auto& pt : pts) pt->start_raytracing();          // every GPU works on its shard at the same time
    for (auto& pt : pts)
      while (!pt->is_done()) std::this_thread::sleep_for(std::chrono::milliseconds(5));   // application.cpp:1984
    if (gpus > 1) {
      std::vector<b2rt_comm*> comms(gpus, nullptr);
      std::vector<b2rt_renderer*> hs;
      for (auto& pt : pts) hs.push_back(pt->handle());
      b2rt_shim::check(b2rt_comm_create_all(gpus, nullptr, comms.data()));
      b2rt_shim::check(b2rt_reduce_accum_all(hs.data(), comms.data(), gpus, 0));
      for (b2rt_comm* c : comms) b2rt_comm_destroy(c);
    }
    b2rt_shim::PathTracer& root = *pts[0];
    root.save_image(out);
    if (!exr.empty()) root.save_exr(exr);
    double rays = 0, ms = 0;
    b2rt_stats st{};
    for (auto& pt : pts) { st = pt->stats(); rays += (double)(st.rays_camera + st.rays_bounce + st.rays_shadow); ms = std::max(ms, st.ms_total); }
    std::cout << "rendered " << scene << " " << w << "x" << h << " " << ns_aa << " spp depth " << depth << " on " << gpus << " GPU(s): " << ms
              << " ms, " << rays / ms / 1e3 << " Mrays/s, BVH " << st.bvh_nodes << " nodes / " << st.bvh_subtrees
              << " subtrees / " << st.bvh_levels << " levels -> " << out << "\n";
  } catch (const std::exception& e) {
    std::cerr << e.what() << "\n";
    return 1;
  }
  return 0;
}
