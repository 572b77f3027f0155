// Renderer handle behind b2rt_renderer (see render.cu)
#pragma once
#include <cuda_runtime.h>

#include "traverse.cuh"

namespace b2rt {

// one wave = a pixel range x a run of samples, traced to full depth before the next wave starts
struct WaveParams {
  uint32_t pix0, n_pix;       // pixel range of this wave
  uint32_t spp;               // samples per pixel in this wave
  uint32_t sample0;           // index of the wave's first sample, relative to *sample_base
  const uint32_t* sample_base;  // device word: the frame's first global sample (cfg.sample_first); behind a pointer so
                              // that a captured frame (CUDA graph) can be replayed for the next samples
  uint32_t sample_stride;
  uint32_t width, height;
  uint32_t jitter;            // 0 -> pixel centre
  uint32_t k0, k1;            // Philox key
  float eps;
  uint32_t max_depth, ns_area_light, S;  // S = shadow rays per interaction
};

struct Renderer {
  b2rt_config cfg{};
  int device = -1;
  cudaStream_t stream = nullptr, own_stream = nullptr;
  cudaStream_t stream_cancel = nullptr;   // b2rt_stop raises the cancel flag from here while `stream` is busy
  cudaEvent_t ev_start = nullptr, ev_done = nullptr;
  // scene
  DeviceBVH dbvh;
  Tracer tracer;
  // second scheduler + stream: the shadow-ray trace of bounce b runs next to the closest-hit trace of bounce b + 1
  Tracer tracer2;
  cudaStream_t stream2 = nullptr;
  std::vector<cudaEvent_t> ev_sync;     // [2 * MAX_DEPTH] shade(b) done / resolve(b) done
  bool overlap = false;
  void* d_prim_geom = nullptr; float* d_tri_normals = nullptr; float* d_tri_normals_buf = nullptr; uint32_t* d_prim_material = nullptr;
  b2rt_material* d_materials = nullptr; b2rt_light* d_lights = nullptr; float* d_light_area = nullptr;
  std::vector<b2rt_light> lights_host;
  float* d_env = nullptr; uint32_t env_w = 0, env_h = 0;   // environment map (b2rt_set_envmap)
  size_t cap_prims = 0, cap_normals = 0, cap_mats = 0, cap_lights = 0;   // grow-only device array capacities (elements)
  uint32_t n_tris = 0, n_lights = 0, n_wide_nodes = 0, shadow_per_hit = 0;
  double build_ms = 0;
  bool have_scene = false, have_camera = false, running = false, bvh_stale = false, have_glossy = false;
  b2rt_camera cam{};
  // frame
  uint32_t width = 0, height = 0;
  void* accum = nullptr; void* img_a = nullptr; void* img_b = nullptr; uint32_t* ldr = nullptr;
  void* resolved = nullptr;
  float* host_image = nullptr; size_t host_image_cap = 0;   // page-locked buffer behind b2rt_get_image (floats)
  uint64_t samples_done = 0;
  // the waves of the frame in flight and their outcome (k_wave_end: 1 complete, 2 queue overflow, 3 cancelled)
  struct FrameCtx;
  std::vector<WaveParams> waves;
  uint32_t* wave_status = nullptr; size_t wave_status_cap = 0;
  uint64_t waves_retried = 0, queues_grown = 0;
  // A frame whose launch sequence is identical to the previous one (same scene buffers, camera, waves, schedulers) is
  // captured into a CUDA graph on its second occurrence and replayed from the third on: at the reference's operating point
  // (2 samples of a 512 x 512 frame per render() call) the frame is ~70 launches of a few microseconds each and the
  // launch gaps are most of its time.  B2RT_GRAPH=0 turns it off; per-launch timing (b2rt_set_profiling) never uses it.
  cudaGraphExec_t graph_exec = nullptr;
  uint64_t graph_sig = 0, last_sig = 0;
  bool graph_off = false;
  uint64_t g_launches = 0, g_t1[3] = {0, 0, 0}, g_t2[3] = {0, 0, 0};   // launch counts of the captured frame
  uint64_t graph_replays = 0;
  uint32_t* d_sample_base = nullptr; uint32_t* h_sample_base = nullptr;   // device word + its page-locked source
  uint64_t frame_signature(const FrameCtx& fc) const;
  bool pair_factor_from_env = false;
  uint32_t pair_factor = 4;   // scheduler queue capacity in pushes per ray and level; doubled by wait() after an overflow
  // wave buffers
  uint64_t wave_cap = 0; uint32_t wave_S = 0;
  uint32_t shade_ctas = 0, list_slack = 0;   // k_shade's fixed grid; null entries a list can hold on top of its paths
  void *l_o[2] = {nullptr, nullptr}, *l_d[2] = {nullptr, nullptr};           // dense ray lists (double-buffered by bounce)
  unsigned long long *l_h[2] = {nullptr, nullptr};
  uint32_t *l_slot[2] = {nullptr, nullptr};
  void *thr = nullptr, *rad = nullptr, *s_o = nullptr, *s_d = nullptr, *s_contrib = nullptr;
  unsigned long long *s_hits = nullptr, *totals = nullptr;
  uint32_t *s_q0 = nullptr, *counts = nullptr;
  // stats
  b2rt_stats last{};
  uint64_t launches = 0, cam_rays_enqueued = 0;
  double ms_total = 0, ms_traverse_acc = 0;

  int set_device();
  int set_stream(cudaStream_t s);
  int create(const b2rt_config* c);
  void destroy();
  void release_scene();
  void release_wave();
  int set_scene(const b2rt_scene_desc* d);
  int set_camera(const b2rt_camera* c);
  int set_envmap(const float* rgb, uint32_t w, uint32_t h);
  int set_frame_size(uint32_t w, uint32_t h);
  int clear();
  int ensure_wave();
  int start();
  int make_frame_ctx(FrameCtx* fc);
  int enqueue_wave(const FrameCtx& fc, const WaveParams& wp, uint32_t status_index);
  int retry_wave(const FrameCtx& fc, const WaveParams& wp, int depth);
  int is_done();
  int wait();
  int stop();
  int resolve(bool want_ldr);
  void fill_stats(b2rt_stats* out) const;
};

}  // namespace b2rt

// the C handle is the Renderer (api.cu, comm.cu)
struct b2rt_renderer { b2rt::Renderer r; };
