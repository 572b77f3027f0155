// Header-only C++ shims with the reference's class shapes, implemented ONLY with the C ABI (include/b2rt.h).
// A Scotty3D / CudaScotty maintainer swaps the class bodies for these calls.
//
//   b2rt_shim::PathTracer    <-> class PathTracer               src/pathtracer.h:51-257
//   b2rt_shim::BVHAccel      <-> class BVHAccel                 src/bvh.h:99-191
//   b2rt_shim::CudaRenderer  <-> class cutracer::CudaRenderer   src/cudaRenderer.h:173-272
//
// Differences that are deliberate: scenes are flat b2rt_scene_desc arrays instead of StaticScene object
// graphs; errors throw std::runtime_error(b2rt_last_error()) instead of exit(); BVH queries are batched.
#pragma once
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#include "../../include/b2rt.h"

namespace b2rt_shim {

inline void check(int rc) {
  if (rc < 0) throw std::runtime_error(std::string("b2rt: ") + b2rt_last_error());
}

// Scene file holder (b2rt_scene_load / b2rt_load_dae)
class SceneFile {
 public:
  explicit SceneFile(const std::string& path) {
    const bool dae = path.size() > 4 && path.compare(path.size() - 4, 4, ".dae") == 0;
    check(dae ? b2rt_load_dae(path.c_str(), &f_) : b2rt_scene_load(path.c_str(), &f_));
  }
  ~SceneFile() { b2rt_scene_free(f_); }
  SceneFile(const SceneFile&) = delete;
  SceneFile& operator=(const SceneFile&) = delete;
  const b2rt_scene_desc* desc() const { return &f_->desc; }
  const b2rt_scene_file* file() const { return f_; }
  // Application::load camera placement for a given frame size (src/application.cpp:395-408)
  b2rt_camera camera(uint32_t w, uint32_t h) const {
    b2rt_camera c;
    check(b2rt_camera_place(f_->bbox, f_->cam_dir, f_->cam_hfov_deg, f_->cam_vfov_deg, w, h, &c));
    return c;
  }

 private:
  b2rt_scene_file* f_ = nullptr;
};

struct Ray {        // src/ray.h
  float o[3], d[3], min_t, max_t;
};
struct Intersection {  // src/intersection.h:21-32 (t, primitive); n / bsdf are looked up by primitive id
  float t;
  uint32_t primitive;  // 0xFFFFFFFF = no hit
};

class BVHAccel {
 public:
  // BVHAccel(const std::vector<Primitive*>&, size_t max_leaf_size = 4)   src/bvh.h:110-112
  // build_on_device: construct the tree with the CUDA builder (b2rt_bvh_build_device) instead of on the host cores
  explicit BVHAccel(const b2rt_scene_desc* prims, size_t max_leaf_size = 4, uint32_t width = 4, bool build_on_device = false) {
    check(build_on_device ? b2rt_bvh_build_device(prims, (uint32_t)max_leaf_size, width, 0, -1, &h_)
                          : b2rt_bvh_build(prims, (uint32_t)max_leaf_size, width, 0, -1, &h_));
  }
  // distance slices of the batch traversal (b2rt_bvh_set_slicing): > 0 first slice length, 0 off, < 0 automatic
  void set_slicing(float first_slice, float growth = 4.f, int passes = 4) { check(b2rt_bvh_set_slicing(h_, first_slice, growth, passes)); }
  ~BVHAccel() { b2rt_bvh_destroy(h_); }
  BVHAccel(const BVHAccel&) = delete;
  BVHAccel& operator=(const BVHAccel&) = delete;
  // get_bbox()   src/bvh.h:120-126
  void get_bbox(float out6[6]) const { check(b2rt_bvh_get_bbox(h_, out6)); }
  // bool intersect(const Ray&, Intersection*) for a batch   src/bvh.h:150-163
  std::vector<Intersection> intersect(const std::vector<Ray>& rays) const {
    const size_t n = rays.size();
    std::vector<float> o(3 * n), d(3 * n), t0(n), t1(n), t(n);
    std::vector<uint32_t> p(n);
    for (size_t i = 0; i < n; ++i) {
      memcpy(&o[3 * i], rays[i].o, 12); memcpy(&d[3 * i], rays[i].d, 12);
      t0[i] = rays[i].min_t; t1[i] = rays[i].max_t;
    }
    check(b2rt_bvh_intersect(h_, o.data(), d.data(), t0.data(), t1.data(), n, t.data(), p.data()));
    std::vector<Intersection> out(n);
    for (size_t i = 0; i < n; ++i) { out[i].t = t[i]; out[i].primitive = p[i]; }
    return out;
  }
  // bool intersect(const Ray&) for a batch   src/bvh.h:139-148
  std::vector<uint8_t> intersect_any(const std::vector<Ray>& rays) const {
    const size_t n = rays.size();
    std::vector<float> o(3 * n), d(3 * n), t0(n), t1(n);
    std::vector<uint8_t> occ(n);
    for (size_t i = 0; i < n; ++i) {
      memcpy(&o[3 * i], rays[i].o, 12); memcpy(&d[3 * i], rays[i].d, 12);
      t0[i] = rays[i].min_t; t1[i] = rays[i].max_t;
    }
    check(b2rt_bvh_occluded(h_, o.data(), d.data(), t0.data(), t1.data(), n, occ.data()));
    return occ;
  }
  b2rt_bvh* handle() const { return h_; }

 private:
  b2rt_bvh* h_ = nullptr;
};

class PathTracer {
 public:
  enum State { INIT, READY, RENDERING, DONE };   // src/pathtracer.h:196-202 (VISUALIZE is GUI-only)

  // PathTracer(ns_aa, max_ray_depth, ns_area_light, ns_diff, ns_glsy, ns_refr, num_threads, envmap)
  // src/pathtracer.h:57-60.  ns_diff/ns_glsy/ns_refr/num_threads/envmap are accepted and ignored (the
  // reference ignores the first three too; threads are CUDA's business here).
  explicit PathTracer(size_t ns_aa = 1, size_t max_ray_depth = 4, size_t ns_area_light = 1, size_t = 1, size_t = 1,
                      size_t = 1, size_t = 1, void* = nullptr) {
    memset(&cfg_, 0, sizeof cfg_);
    cfg_.ns_aa = (uint32_t)ns_aa; cfg_.max_ray_depth = (uint32_t)max_ray_depth; cfg_.ns_area_light = (uint32_t)ns_area_light;
    cfg_.device = -1; cfg_.sample_stride = 1;
    check(b2rt_create(&cfg_, &h_));
  }
  // every knob of the C ABI at once (device ordinal, sample shard of a multi-GPU job, BVH parameters, ...)
  explicit PathTracer(const b2rt_config& cfg) : cfg_(cfg) { check(b2rt_create(&cfg_, &h_)); }
  ~PathTracer() { b2rt_destroy(h_); }
  PathTracer(const PathTracer&) = delete;
  PathTracer& operator=(const PathTracer&) = delete;

  void set_scene(const b2rt_scene_desc* scene) {            // src/pathtracer.cpp:71-92 (+ build_accel)
    check(b2rt_set_scene(h_, scene)); have_scene_ = true; to_ready();
  }
  void set_camera(const b2rt_camera* camera) {              // src/pathtracer.cpp:94-103
    check(b2rt_set_camera(h_, camera)); have_camera_ = true; to_ready();
  }
  void set_frame_size(size_t width, size_t height) {        // src/pathtracer.cpp:105-114
    check(b2rt_set_frame_size(h_, (uint32_t)width, (uint32_t)height));
    w_ = (uint32_t)width; h_px_ = (uint32_t)height; to_ready();
  }
  void start_raytracing() {                                 // src/pathtracer.cpp:183-213: only from READY
    if (state_ != READY && state_ != DONE) return;
    check(b2rt_clear(h_));                                  // the reference clears its sample / frame buffers here
    check(b2rt_start(h_)); state_ = RENDERING;
  }
  bool is_done() {                                          // src/pathtracer.cpp:572-575
    if (state_ == RENDERING) { int r = b2rt_is_done(h_); check(r); if (r == 1) state_ = DONE; }
    return state_ == DONE;
  }
  void stop() {                                             // src/pathtracer.cpp:116-139
    check(b2rt_stop(h_)); if (state_ == RENDERING || state_ == DONE) state_ = READY;
  }
  void clear() { check(b2rt_clear(h_)); }
  void increase_area_light_sample_count() { cfg_.ns_area_light *= 2; check(b2rt_set_config(h_, &cfg_)); }   // :560-564
  void decrease_area_light_sample_count() { if (cfg_.ns_area_light > 1) cfg_.ns_area_light /= 2; check(b2rt_set_config(h_, &cfg_)); }
  void key_press(int key) {                                 // src/pathtracer.cpp:357-367
    if (key == '[') { if (cfg_.ns_aa > 1) cfg_.ns_aa /= 2; } else if (key == ']') cfg_.ns_aa *= 2; else return;
    check(b2rt_set_config(h_, &cfg_));
  }
  State state() const { return state_; }
  // HDRImageBuffer: Spectrum per pixel, index x + y*w (src/image.h:114-118)
  std::vector<float> hdr() { std::vector<float> v((size_t)w_ * h_px_ * 3); check(b2rt_read_hdr(h_, v.data(), v.size())); return v; }
  // ImageBuffer: RGBA8 via toColor (src/image.h:49-58,173-188)
  std::vector<uint32_t> frame() { std::vector<uint32_t> v((size_t)w_ * h_px_); check(b2rt_read_ldr(h_, v.data(), v.size())); return v; }
  b2rt_stats stats() { b2rt_stats s; check(b2rt_get_stats(h_, &s)); return s; }
  // save_image: PNG, vertically flipped like src/pathtracer.cpp:577-591 (b2rt_write_png); save_exr: the HDR frame
  void save_image(const std::string& filename) { check(b2rt_write_png(h_, filename.c_str())); }
  void save_exr(const std::string& filename) { check(b2rt_write_exr(h_, filename.c_str())); }
  void wait() { check(b2rt_wait(h_)); if (state_ == RENDERING) state_ = DONE; }
  b2rt_renderer* handle() const { return h_; }

 private:
  void to_ready() {
    if (state_ == INIT) { if (have_scene_ && have_camera_ && w_) state_ = READY; }
    else state_ = READY;
  }
  b2rt_renderer* h_ = nullptr;
  b2rt_config cfg_;
  State state_ = INIT;
  bool have_scene_ = false, have_camera_ = false;
  uint32_t w_ = 0, h_px_ = 0;
};

// Progressive renderer with the CudaRenderer call sequence of src/cudaMain.cpp:88-99 / src/display.cpp:145-173
class CudaRenderer {
 public:
  // SAMPLES_PER_PIXEL 2, three surface hits, POST_PROCESS_THRESHOLD 32   src/cudaRenderer.h:70-73
  explicit CudaRenderer(uint32_t samples_per_frame = 2, uint32_t max_ray_depth = 3, uint32_t ns_area_light = 2,
                        uint32_t median_threshold = 32) {
    memset(&cfg_, 0, sizeof cfg_);
    cfg_.ns_aa = samples_per_frame; cfg_.max_ray_depth = max_ray_depth; cfg_.ns_area_light = ns_area_light;
    cfg_.median_threshold = median_threshold; cfg_.device = -1; cfg_.sample_stride = 1;
    check(b2rt_create(&cfg_, &h_));
  }
  ~CudaRenderer() { b2rt_destroy(h_); delete scene_; }
  void allocOutputImage(int width, int height) { w_ = width; hgt_ = height; check(b2rt_set_frame_size(h_, width, height)); }
  void loadScene(const std::string& name) { delete scene_; scene_ = new SceneFile(name); check(b2rt_set_scene(h_, scene_->desc())); }
  void setup() { b2rt_camera c = scene_->camera(w_, hgt_); check(b2rt_set_camera(h_, &c)); frames_ = 0; }
  void setViewpoint(const b2rt_camera& cam) { check(b2rt_set_camera(h_, &cam)); frames_ = 0; }   // resets accumulation, :1866-1869
  // setViewpoint(Vector3D origin, Vector3D lookAt), src/cudaRenderer.cu:1845-1870, with the reference's camera basis
  // (left = (0,1,0) x -lookAt, up = left x -lookAt, :1592-1599) and its fixed frustum k = (u - .5, -(v - .5), 1) (:347)
  void setViewpoint(const float origin[3], const float lookAt[3]) {
    memcpy(c_origin, origin, 12); memcpy(c_lookAt, lookAt, 12);
    b2rt_camera c;
    check(b2rt_camera_look_at(origin, lookAt, 0.f, &c));
    setViewpoint(c);
  }
  float c_origin[3] = {0, 0, 0}, c_lookAt[3] = {0, 0, -1};   // public like the reference's (src/cudaRenderer.h:250-254)
  void clearImage() { check(b2rt_clear(h_)); frames_ = 0; }
  void render() {                                   // renderAccumulate, src/cudaRenderer.cu:2419-2457
    cfg_.sample_first = frames_ * cfg_.ns_aa;
    check(b2rt_set_config(h_, &cfg_));
    check(b2rt_render(h_));
    ++frames_;
  }
  const float* getImage() {                         // float4 RGBA, row-major x + y*w (NOT the reference's x*H + y)
    const float* img = nullptr;                     // renderer-owned, valid until the next call (like the reference)
    check(b2rt_get_image(h_, &img, nullptr));
    return img;
  }
  b2rt_renderer* handle() const { return h_; }

 private:
  b2rt_renderer* h_ = nullptr;
  b2rt_config cfg_;
  SceneFile* scene_ = nullptr;
  int w_ = 0, hgt_ = 0;
  uint32_t frames_ = 0;
};

}  // namespace b2rt_shim
