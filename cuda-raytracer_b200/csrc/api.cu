// C ABI of the b2rt library (include/b2rt.h).  Thin: argument checks, handle management, host<->device
// copies of caller buffers.  Every function documents the reference interface it replaces in b2rt.h.
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "render.cuh"
#include "rt_device.cuh"

namespace b2rt {
static thread_local std::string g_error;
void set_error(const std::string& msg) { g_error = msg; }
}  // namespace b2rt

using namespace b2rt;

struct b2rt_bvh {
  int device = 0;
  DeviceBVH dbvh;
  Tracer tracer;
  WideBVH host_meta;     // blob released after upload; keeps levels / counts
  cudaStream_t stream = nullptr;
  uint64_t ray_cap = 0;
  uint32_t pair_factor = 0;
  float4 *ray_o = nullptr, *ray_d = nullptr;
  unsigned long long* hits = nullptr;
  uint32_t* n_dev = nullptr;
  uint32_t max_leaf = 4, treelet_budget = 0;   // as given to the builder (b2rt_bvh_validate)
  b2rt_stats last{};
};

namespace {

__global__ void k_pack_rays(const float* __restrict__ org, const float* __restrict__ dir, const float* __restrict__ tmin,
                            const float* __restrict__ tmax, uint32_t n, float4* ro, float4* rd, unsigned long long* hits) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  // The traversal orders distances as unsigned integers (packed (t, prim) words, stack / sort keys), which is the
  // float order only for t >= 0: t_min is clamped to >= 0 and a negative or NaN t_max becomes 0 (b2rt.h documents the
  // t >= 0 contract).
  const float t0 = tmin[i] > 0.f ? tmin[i] : 0.f, t1 = tmax[i] >= 0.f ? tmax[i] : 0.f;
  ro[i] = make_float4(org[3 * i], org[3 * i + 1], org[3 * i + 2], t0);
  rd[i] = make_float4(dir[3 * i], dir[3 * i + 1], dir[3 * i + 2], t1);
  hits[i] = pack_hit(t1, 0xFFFFFFFFu);
}
__global__ void k_unpack_hits(const unsigned long long* __restrict__ hits, uint32_t n, float* t, uint32_t* prim, uint8_t* occ) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  unsigned long long h = hits[i];
  uint32_t p = (uint32_t)h;
  if (occ) occ[i] = p != 0xFFFFFFFFu;
  if (t) t[i] = p == 0xFFFFFFFFu ? __builtin_huge_valf() : __uint_as_float((uint32_t)(h >> 32));
  if (prim) prim[i] = p;
}
// synthetic rays for kernel-only timing: mode 0 coherent (pinhole toward the box), mode 1 incoherent
__global__ void k_gen_rays(uint32_t n, int mode, uint32_t k0, uint32_t k1, float3 lo, float3 hi, float4* ro, float4* rd,
                           unsigned long long* hits) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float3 c = make_float3(0.5f * (lo.x + hi.x), 0.5f * (lo.y + hi.y), 0.5f * (lo.z + hi.z));
  const float3 e = make_float3(hi.x - lo.x, hi.y - lo.y, hi.z - lo.z);
  const float R = 0.5f * sqrtf(e.x * e.x + e.y * e.y + e.z * e.z);
  f3 o, d;
  if (mode == 0) {
    // camera at c + (0,0,2.2R), image plane side^2 pixels, 8x4-pixel tiles per warp for coherence
    const uint32_t side = (uint32_t)ceilf(sqrtf((float)n));
    const uint32_t tile = i / 32, in = i % 32;
    const uint32_t tiles_x = (side + 7) / 8;
    const uint32_t px = (tile % tiles_x) * 8 + (in % 8), py = (tile / tiles_x) * 4 + (in / 8);
    const uint4 r = philox4x32_10(i, 0, 0, 7, k0, k1);
    const float sx = ((float)px + u01(r.x)) / (float)side, sy = ((float)py + u01(r.y)) / (float)(side);
    o = mk3(c.x, c.y, c.z + 2.2f * R);
    d = normalize3(mk3((2.f * sx - 1.f) * 0.6f, (2.f * sy - 1.f) * 0.6f, -1.f));
  } else {
    const uint4 r = philox4x32_10(i, 1, 0, 7, k0, k1);
    const uint4 q = philox4x32_10(i, 2, 0, 7, k0, k1);
    o = mk3(lo.x + e.x * u01(r.x), lo.y + e.y * u01(r.y), lo.z + e.z * u01(r.z));
    const float z = 1.f - 2.f * u01(q.x);
    float s, cs;
    sincos2pi(u01(q.y), &s, &cs);
    const float rr = sqrtf(fmaxf(0.f, 1.f - z * z));
    d = mk3(rr * cs, rr * s, z);
  }
  ro[i] = make_float4(o.x, o.y, o.z, 0.f);
  rd[i] = make_float4(d.x, d.y, d.z, __builtin_huge_valf());
  hits[i] = pack_hit(__builtin_huge_valf(), 0xFFFFFFFFu);
}
__global__ void k_reset_hits(uint32_t n, const float4* rd, unsigned long long* hits) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) hits[i] = pack_hit(rd[i].w, 0xFFFFFFFFu);
}
__global__ void k_count_hits(uint32_t n, const unsigned long long* hits, unsigned long long* out) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  bool h = i < n && (uint32_t)hits[i] != 0xFFFFFFFFu;
  uint32_t m = __ballot_sync(0xffffffffu, h);
  if ((threadIdx.x & 31) == 0 && m) atomicAdd(out, (unsigned long long)__popc(m));
}

int ensure_rays(b2rt_bvh* b, uint64_t n, uint32_t pair_factor = 6) {
  if (b->ray_cap >= n && b->pair_factor >= pair_factor) return B2RT_OK;
  if (b->ray_o) { cudaFree(b->ray_o); cudaFree(b->ray_d); cudaFree(b->hits); b->ray_o = b->ray_d = nullptr; b->hits = nullptr; }
  b->tracer.release();
  b->ray_cap = 0;
  B2RT_CUDA_OK(cudaMalloc(&b->ray_o, (n + 16) * sizeof(float4)));
  B2RT_CUDA_OK(cudaMalloc(&b->ray_d, (n + 16) * sizeof(float4)));
  B2RT_CUDA_OK(cudaMalloc(&b->hits, (n + 16) * 8));
  int rc = b->tracer.init(b->dbvh, n, pair_factor);
  if (rc) return rc;
  b->ray_cap = n; b->pair_factor = pair_factor;
  return B2RT_OK;
}

// distance slicing of the batch traversal (b2rt_bvh_set_slicing)
void configure_slicing(b2rt_bvh* b, float first, float growth, int passes) {
  Tracer& T = b->tracer;
  const float* bb = b->host_meta.bbox;
  for (int k = 0; k < 6; ++k) T.slice_bbox[k] = bb[k];
  const float ex = bb[3] - bb[0], ey = bb[4] - bb[1], ez = bb[5] - bb[2];
  const float diag = std::sqrt(ex * ex + ey * ey + ez * ez);
  if (first < 0.f) {
    const float want = 2.f * b->host_meta.mean_free_path;   // profiles/r01_sweep_slices.txt
    first = (b->dbvh.n_levels >= 3 && want > 0.f && want < 0.125f * diag) ? want : 0.f;
  }
  T.slice_first = first;
  T.slice_growth = growth > 1.f ? growth : 4.f;
  T.slice_passes = passes >= 2 ? std::min(passes, 16) : 4;
}

int has_device() {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

// closest / any-hit on host rays in batches (split on queue overflow)
int intersect_host(b2rt_bvh* b, const float* org, const float* dir, const float* tmin, const float* tmax, uint64_t n,
                   float* hit_t, uint32_t* hit_prim, uint8_t* occluded, bool any_hit) {
  if (!b || (n && (!org || !dir || !tmin || !tmax))) { set_error("null argument"); return B2RT_ERR_INVALID; }
  B2RT_CUDA_OK(cudaSetDevice(b->device));
  uint64_t batch = std::min<uint64_t>(n, 1u << 22);
  if (batch == 0) return B2RT_OK;
  float *d_org = nullptr, *d_dir = nullptr, *d_tmin = nullptr, *d_tmax = nullptr, *d_t = nullptr;
  uint32_t* d_prim = nullptr; uint8_t* d_occ = nullptr;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  auto cleanup = [&]() {
    cudaFree(d_org); cudaFree(d_dir); cudaFree(d_tmin); cudaFree(d_tmax); cudaFree(d_t); cudaFree(d_prim); cudaFree(d_occ);
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
  };
  int rc = ensure_rays(b, batch);
  if (rc) return rc;
  if (cudaMalloc(&d_org, batch * 12) || cudaMalloc(&d_dir, batch * 12) || cudaMalloc(&d_tmin, batch * 4) ||
      cudaMalloc(&d_tmax, batch * 4) || cudaMalloc(&d_t, batch * 4) || cudaMalloc(&d_prim, batch * 4) || cudaMalloc(&d_occ, batch)) {
    cleanup(); set_error("cudaMalloc failed for ray staging"); cudaGetLastError(); return B2RT_ERR_OOM;
  }
  cudaStream_t s = b->stream;
  b->tracer.launches = 0;
  if (cudaEventCreate(&e0) != cudaSuccess || cudaEventCreate(&e1) != cudaSuccess) {
    cleanup(); set_error("cudaEventCreate failed"); cudaGetLastError(); return B2RT_ERR_CUDA;
  }
  double ms_sum = 0;
  uint64_t done = 0;
  while (done < n) {
    uint64_t m = std::min<uint64_t>(batch, n - done);
    for (;;) {
      cudaMemcpyAsync(d_org, org + 3 * done, m * 12, cudaMemcpyHostToDevice, s);
      cudaMemcpyAsync(d_dir, dir + 3 * done, m * 12, cudaMemcpyHostToDevice, s);
      cudaMemcpyAsync(d_tmin, tmin + done, m * 4, cudaMemcpyHostToDevice, s);
      cudaMemcpyAsync(d_tmax, tmax + done, m * 4, cudaMemcpyHostToDevice, s);
      uint32_t m32 = (uint32_t)m;
      cudaMemcpyAsync(b->n_dev, &m32, 4, cudaMemcpyHostToDevice, s);
      cudaEventRecord(e0, s);
      k_pack_rays<<<(m32 + 255) / 256, 256, 0, s>>>(d_org, d_dir, d_tmin, d_tmax, m32, b->ray_o, b->ray_d, b->hits);
      rc = b->tracer.trace_sliced(s, b->ray_o, b->ray_d, b->hits, b->n_dev, m, any_hit);
      if (rc) { cleanup(); return rc; }
      k_unpack_hits<<<(m32 + 255) / 256, 256, 0, s>>>(b->hits, m32, d_t, d_prim, d_occ);
      cudaEventRecord(e1, s);
      bool ovf = false;
      rc = b->tracer.check_overflow(s, &ovf);
      if (rc) { cleanup(); return rc; }
      if (!ovf) break;
      if (m <= 1024) { cleanup(); set_error("ray queue overflow on a minimal batch"); return B2RT_ERR_OVERFLOW; }
      m = m / 2;   // split the batch and retry
    }
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1); ms_sum += ms;
    if (hit_t) cudaMemcpyAsync(hit_t + done, d_t, m * 4, cudaMemcpyDeviceToHost, s);
    if (hit_prim) cudaMemcpyAsync(hit_prim + done, d_prim, m * 4, cudaMemcpyDeviceToHost, s);
    if (occluded) cudaMemcpyAsync(occluded + done, d_occ, m, cudaMemcpyDeviceToHost, s);
    cudaError_t e = cudaStreamSynchronize(s);
    if (e != cudaSuccess) { cleanup(); set_error(std::string("intersect: ") + cudaGetErrorString(e)); return B2RT_ERR_CUDA; }
    done += m;
  }
  cleanup();
  TraceCounters tc;
  b->tracer.read_counters(&tc, nullptr);
  b->last.node_visits = tc.node_visits; b->last.leaf_prim_tests = tc.prim_tests; b->last.subtree_visits = tc.subtree_visits;
  b->last.queue_pushes = tc.pushes; b->last.staged_bytes = tc.staged_bytes; b->last.hit_updates = tc.hit_updates;
  b->last.kernel_launches = b->tracer.launches; b->last.ms_total = ms_sum; b->last.ms_traverse = ms_sum;
  b->tracer.reset_counters(b->stream);
  cudaStreamSynchronize(b->stream);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { set_error(std::string("intersect: ") + cudaGetErrorString(e)); return B2RT_ERR_CUDA; }
  return B2RT_OK;
}

}  // namespace

extern "C" {

const char* b2rt_last_error(void) { return g_error.c_str(); }
int b2rt_abi_version(void) { return B2RT_ABI_VERSION; }
int b2rt_device_count(void) { return has_device(); }

// every failure after the handle exists goes through b2rt_bvh_destroy (stream, device buffers, the handle itself)
#define B2RT_BVH_OK(call)                                                                            \
  do {                                                                                               \
    cudaError_t e__ = (call);                                                                        \
    if (e__ != cudaSuccess) {                                                                        \
      set_error(std::string(#call) + ": " + cudaGetErrorString(e__));                                \
      b2rt_bvh_destroy(b);                                                                           \
      return e__ == cudaErrorMemoryAllocation ? B2RT_ERR_OOM : B2RT_ERR_CUDA;                        \
    }                                                                                                \
  } while (0)

int b2rt_bvh_build(const b2rt_scene_desc* scene, uint32_t max_leaf_size, uint32_t width, uint32_t treelet_bytes,
                   int32_t device, b2rt_bvh** out) {
  if (!out) { set_error("out is null"); return B2RT_ERR_INVALID; }
  *out = nullptr;
  if (!has_device()) { set_error("no CUDA device available (b2rt has no CPU fallback)"); return B2RT_ERR_NO_DEVICE; }
  HostScene hs;
  int rc = make_host_scene(scene, &hs);
  if (rc) return rc;
  b2rt_bvh* b = new (std::nothrow) b2rt_bvh();
  if (!b) return B2RT_ERR_OOM;
  if (device < 0) cudaGetDevice(&device);
  b->device = device;
  B2RT_BVH_OK(cudaSetDevice(device));
  rc = build_wide_bvh(hs, max_leaf_size, width, treelet_bytes, &b->host_meta);
  if (rc) { b2rt_bvh_destroy(b); return rc; }
  rc = upload_bvh(b->host_meta, &b->dbvh);
  if (rc) { b2rt_bvh_destroy(b); return rc; }
  b->host_meta.blob.clear(); b->host_meta.blob.shrink_to_fit();
  b->max_leaf = max_leaf_size ? max_leaf_size : 4; b->treelet_budget = treelet_bytes;
  B2RT_BVH_OK(cudaStreamCreateWithFlags(&b->stream, cudaStreamNonBlocking));
  B2RT_BVH_OK(cudaMalloc(&b->n_dev, 4));
  b->tracer.bvh = b->dbvh;
  b->tracer.collect_stats = true;
  configure_slicing(b, -1.f, 0.f, 0);
  *out = b;
  return B2RT_OK;
}

int b2rt_bvh_build_device(const b2rt_scene_desc* scene, uint32_t max_leaf_size, uint32_t width, uint32_t treelet_bytes,
                          int32_t device, b2rt_bvh** out) {
  if (!out) { set_error("out is null"); return B2RT_ERR_INVALID; }
  *out = nullptr;
  if (!has_device()) { set_error("no CUDA device available (b2rt has no CPU fallback)"); return B2RT_ERR_NO_DEVICE; }
  b2rt_bvh* b = new (std::nothrow) b2rt_bvh();
  if (!b) return B2RT_ERR_OOM;
  if (device < 0) cudaGetDevice(&device);
  b->device = device;
  B2RT_BVH_OK(cudaSetDevice(device));
  B2RT_BVH_OK(cudaStreamCreateWithFlags(&b->stream, cudaStreamNonBlocking));
  int rc = build_wide_bvh_device(scene, max_leaf_size, width, treelet_bytes, b->stream, &b->dbvh, &b->host_meta);
  if (rc) { b2rt_bvh_destroy(b); return rc; }
  b->max_leaf = max_leaf_size ? max_leaf_size : 3; b->treelet_budget = treelet_bytes;
  B2RT_BVH_OK(cudaMalloc(&b->n_dev, 4));
  b->tracer.bvh = b->dbvh;
  b->tracer.collect_stats = true;
  configure_slicing(b, -1.f, 0.f, 0);
  *out = b;
  return B2RT_OK;
}
#undef B2RT_BVH_OK

int b2rt_bvh_validate(b2rt_bvh* b, const b2rt_scene_desc* scene, uint64_t out8[8]) {
  if (!b || !scene) { set_error("null argument"); return B2RT_ERR_INVALID; }
  HostScene hs;
  int rc = make_host_scene(scene, &hs);
  if (rc) return rc;
  B2RT_CUDA_OK(cudaSetDevice(b->device));
  WideBVH wb;
  wb.width = b->dbvh.width; wb.n_levels = b->dbvh.n_levels; wb.max_treelet_bytes = b->dbvh.max_treelet_bytes;
  wb.levels.assign(b->dbvh.levels, b->dbvh.levels + b->dbvh.n_levels);
  wb.treelets.resize(b->dbvh.n_treelets);
  wb.blob.resize(b->dbvh.blob_bytes);
  B2RT_CUDA_OK(cudaStreamSynchronize(b->stream));
  B2RT_CUDA_OK(cudaMemcpy(wb.treelets.data(), b->dbvh.treelets, wb.treelets.size() * sizeof(TreeletDesc), cudaMemcpyDeviceToHost));
  B2RT_CUDA_OK(cudaMemcpy(wb.blob.data(), b->dbvh.blob, wb.blob.size(), cudaMemcpyDeviceToHost));
  return validate_wide_bvh(hs, wb, b->max_leaf, b->treelet_budget, out8);
}

int b2rt_bvh_set_slicing(b2rt_bvh* b, float first_slice, float growth, int32_t passes) {
  if (!b) { set_error("null argument"); return B2RT_ERR_INVALID; }
  if (!(first_slice == first_slice)) { set_error("first_slice is NaN"); return B2RT_ERR_INVALID; }
  configure_slicing(b, first_slice, growth, passes);
  return B2RT_OK;
}

int b2rt_bvh_intersect(b2rt_bvh* bvh, const float* org, const float* dir, const float* tmin, const float* tmax, uint64_t n,
                       float* hit_t, uint32_t* hit_prim) {
  return intersect_host(bvh, org, dir, tmin, tmax, n, hit_t, hit_prim, nullptr, false);
}
int b2rt_bvh_occluded(b2rt_bvh* bvh, const float* org, const float* dir, const float* tmin, const float* tmax, uint64_t n,
                      uint8_t* occluded) {
  return intersect_host(bvh, org, dir, tmin, tmax, n, nullptr, nullptr, occluded, true);
}

int b2rt_bvh_bench_rays(b2rt_bvh* b, uint64_t n, int mode, uint64_t seed, int repeats, int any_hit, double* ms_per_repeat,
                        uint64_t* hits_out) {
  if (!b || n == 0 || n > 0x7FFFFFFFull || repeats < 1) { set_error("bad argument"); return B2RT_ERR_INVALID; }
  B2RT_CUDA_OK(cudaSetDevice(b->device));
  const bool keep_stats = b->tracer.collect_stats;
  int rc = ensure_rays(b, n, 48);
  if (rc) return rc;
  cudaStream_t s = b->stream;
  const uint32_t n32 = (uint32_t)n;
  B2RT_CUDA_OK(cudaMemcpyAsync(b->n_dev, &n32, 4, cudaMemcpyHostToDevice, s));
  const float* bb = b->host_meta.bbox;
  k_gen_rays<<<(n32 + 255) / 256, 256, 0, s>>>(n32, mode, (uint32_t)seed, (uint32_t)(seed >> 32), make_float3(bb[0], bb[1], bb[2]),
                                               make_float3(bb[3], bb[4], bb[5]), b->ray_o, b->ray_d, b->hits);
  // The batch is traced in sub-batches of `sub` rays; a queue overflow (very incoherent rays on a deep subtree
  // graph push one ray to dozens of subtrees per level) halves `sub` and starts over.
  uint32_t sub = n32;
  auto trace_all = [&]() -> int {
    for (uint32_t first = 0; first < n32; first += sub) {
      const uint32_t m = std::min(sub, n32 - first);
      B2RT_CUDA_OK(cudaMemcpyAsync(b->n_dev, &m, 4, cudaMemcpyHostToDevice, s));
      int r2 = b->tracer.trace_sliced(s, b->ray_o + first, b->ray_d + first, b->hits + first, b->n_dev, m, any_hit != 0);
      if (r2) return r2;
    }
    return B2RT_OK;
  };
  // one untimed pass with statistics, then `repeats` timed passes without
  TraceCounters tc;
  for (;;) {
    b->tracer.launches = 0;
    { int r3 = b->tracer.reset_counters(s); if (r3) return r3; }
    b->tracer.collect_stats = true;
    rc = trace_all();
    b->tracer.collect_stats = false;
    if (rc) { b->tracer.collect_stats = keep_stats; return rc; }
    bool ovf = false;
    rc = b->tracer.check_overflow(s, &ovf);
    if (rc) { b->tracer.collect_stats = keep_stats; return rc; }
    if (!ovf) break;
    if (sub <= 65536) { b->tracer.collect_stats = keep_stats; set_error("ray queue overflow on a 64 Ki-ray batch"); return B2RT_ERR_OVERFLOW; }
    sub = ((sub / 2) + 3u) & ~3u;
    k_reset_hits<<<(n32 + 255) / 256, 256, 0, s>>>(n32, b->ray_d, b->hits);
  }
  { int r3 = b->tracer.read_counters(&tc, nullptr); if (r3) return r3; }
  struct EventPair {
    cudaEvent_t a = nullptr, b = nullptr;
    ~EventPair() { if (a) cudaEventDestroy(a); if (b) cudaEventDestroy(b); }
  } ev;
  struct StatsRestore {   // the timed passes run without counters; the caller's setting comes back on every exit path
    Tracer& t; bool keep;
    ~StatsRestore() { t.collect_stats = keep; }
  } restore{b->tracer, keep_stats};
  B2RT_CUDA_OK(cudaEventCreate(&ev.a));
  B2RT_CUDA_OK(cudaEventCreate(&ev.b));
  double total = 0;
  for (int r = 0; r < repeats; ++r) {
    k_reset_hits<<<(n32 + 255) / 256, 256, 0, s>>>(n32, b->ray_d, b->hits);
    cudaEventRecord(ev.a, s);
    rc = trace_all();
    cudaEventRecord(ev.b, s);
    if (rc) return rc;
    B2RT_CUDA_OK(cudaEventSynchronize(ev.b));
    float ms = 0; cudaEventElapsedTime(&ms, ev.a, ev.b); total += ms;
  }
  if (ms_per_repeat) *ms_per_repeat = total / repeats;
  struct DevWord {
    unsigned long long* p = nullptr;
    ~DevWord() { cudaFree(p); }
  } cnt;
  B2RT_CUDA_OK(cudaMalloc(&cnt.p, 8));
  B2RT_CUDA_OK(cudaMemsetAsync(cnt.p, 0, 8, s));
  k_count_hits<<<(n32 + 255) / 256, 256, 0, s>>>(n32, b->hits, cnt.p);
  unsigned long long hc = 0;
  B2RT_CUDA_OK(cudaMemcpyAsync(&hc, cnt.p, 8, cudaMemcpyDeviceToHost, s));
  B2RT_CUDA_OK(cudaStreamSynchronize(s));
  if (hits_out) *hits_out = hc;
  b->last.node_visits = tc.node_visits; b->last.leaf_prim_tests = tc.prim_tests; b->last.subtree_visits = tc.subtree_visits;
  b->last.queue_pushes = tc.pushes; b->last.staged_bytes = tc.staged_bytes; b->last.hit_updates = tc.hit_updates;
  b->last.ms_total = total / repeats; b->last.ms_traverse = total / repeats;
  b->last.kernel_launches = b->tracer.launches / (uint64_t)(repeats + 1);
  b->last.rays_camera = n;
  return B2RT_OK;
}

int b2rt_bvh_get_stats(b2rt_bvh* b, b2rt_stats* out) {
  if (!b || !out) { set_error("null argument"); return B2RT_ERR_INVALID; }
  *out = b->last;
  out->ms_build = b->host_meta.build_ms;
  out->bvh_nodes = b->host_meta.n_wide_nodes; out->bvh_subtrees = b->dbvh.n_treelets; out->bvh_levels = b->dbvh.n_levels;
  out->bvh_width = b->dbvh.width; out->bvh_bytes = b->dbvh.blob_bytes;
  return B2RT_OK;
}
int b2rt_bvh_get_bbox(b2rt_bvh* b, float* out6) {
  if (!b || !out6) { set_error("null argument"); return B2RT_ERR_INVALID; }
  memcpy(out6, b->host_meta.bbox, 24);
  return B2RT_OK;
}
void b2rt_bvh_destroy(b2rt_bvh* b) {
  if (!b) return;
  cudaSetDevice(b->device);
  if (b->stream) cudaStreamSynchronize(b->stream);
  b->tracer.release();
  free_bvh(&b->dbvh);
  cudaFree(b->ray_o); cudaFree(b->ray_d); cudaFree(b->hits); cudaFree(b->n_dev);
  if (b->stream) cudaStreamDestroy(b->stream);
  delete b;
}

// ---- renderer ----
int b2rt_create(const b2rt_config* cfg, b2rt_renderer** out) {
  if (!cfg || !out) { set_error("null argument"); return B2RT_ERR_INVALID; }
  *out = nullptr;
  b2rt_renderer* h = new (std::nothrow) b2rt_renderer();
  if (!h) return B2RT_ERR_OOM;
  int rc = h->r.create(cfg);
  if (rc) { delete h; return rc; }
  *out = h;
  return B2RT_OK;
}
int b2rt_set_config(b2rt_renderer* h, const b2rt_config* cfg) {
  if (!h || !cfg) { set_error("null argument"); return B2RT_ERR_INVALID; }
  if (h->r.running) { int rc = h->r.wait(); if (rc) return rc; }
  const b2rt_config old = h->r.cfg;
  h->r.cfg = *cfg;
  h->r.cfg.device = old.device;  // the device is fixed at creation
  // BVH parameters only take effect at the next set_scene
  return B2RT_OK;
}
int b2rt_set_scene(b2rt_renderer* h, const b2rt_scene_desc* s) { if (!h) { set_error("null handle"); return B2RT_ERR_INVALID; } return h->r.set_scene(s); }
int b2rt_set_envmap(b2rt_renderer* h, const float* rgb, uint32_t w, uint32_t hh) { if (!h) { set_error("null handle"); return B2RT_ERR_INVALID; } return h->r.set_envmap(rgb, w, hh); }
int b2rt_set_camera(b2rt_renderer* h, const b2rt_camera* c) { if (!h || !c) { set_error("null argument"); return B2RT_ERR_INVALID; } return h->r.set_camera(c); }
int b2rt_set_frame_size(b2rt_renderer* h, uint32_t w, uint32_t hh) { if (!h) { set_error("null handle"); return B2RT_ERR_INVALID; } return h->r.set_frame_size(w, hh); }
int b2rt_start(b2rt_renderer* h) { if (!h) { set_error("null handle"); return B2RT_ERR_INVALID; } return h->r.start(); }
int b2rt_is_done(b2rt_renderer* h) { if (!h) { set_error("null handle"); return B2RT_ERR_INVALID; } return h->r.is_done(); }
int b2rt_wait(b2rt_renderer* h) { if (!h) { set_error("null handle"); return B2RT_ERR_INVALID; } return h->r.wait(); }
int b2rt_stop(b2rt_renderer* h) { if (!h) { set_error("null handle"); return B2RT_ERR_INVALID; } return h->r.stop(); }
int b2rt_clear(b2rt_renderer* h) { if (!h) { set_error("null handle"); return B2RT_ERR_INVALID; } return h->r.clear(); }
int b2rt_render(b2rt_renderer* h) {
  if (!h) { set_error("null handle"); return B2RT_ERR_INVALID; }
  int rc = h->r.start();
  if (rc) return rc;
  return h->r.wait();
}

static int read_common(b2rt_renderer* h, bool ldr) {
  if (!h) { set_error("null handle"); return B2RT_ERR_INVALID; }
  return h->r.resolve(ldr);
}
int b2rt_read_rgba32f(b2rt_renderer* h, float* rgba, size_t n_floats) {
  int rc = read_common(h, false);
  if (rc) return rc;
  const size_t np = (size_t)h->r.width * h->r.height;
  if (!rgba || n_floats < np * 4) { set_error("output buffer too small"); return B2RT_ERR_INVALID; }
  B2RT_CUDA_OK(cudaMemcpyAsync(rgba, h->r.resolved, np * 16, cudaMemcpyDeviceToHost, h->r.stream));
  B2RT_CUDA_OK(cudaStreamSynchronize(h->r.stream));
  return B2RT_OK;
}
int b2rt_get_image(b2rt_renderer* h, const float** rgba, size_t* n_floats) {
  int rc = read_common(h, false);
  if (rc) return rc;
  if (!rgba) { set_error("null argument"); return B2RT_ERR_INVALID; }
  Renderer& R = h->r;
  const size_t np = (size_t)R.width * R.height;
  if (R.host_image_cap < np * 4) {   // grow-only page-locked buffer
    if (R.host_image) cudaFreeHost(R.host_image);
    R.host_image = nullptr; R.host_image_cap = 0;
    B2RT_CUDA_OK(cudaMallocHost(&R.host_image, np * 16));
    R.host_image_cap = np * 4;
  }
  B2RT_CUDA_OK(cudaMemcpyAsync(R.host_image, R.resolved, np * 16, cudaMemcpyDeviceToHost, R.stream));
  B2RT_CUDA_OK(cudaStreamSynchronize(R.stream));
  *rgba = R.host_image;
  if (n_floats) *n_floats = np * 4;
  return B2RT_OK;
}
int b2rt_read_hdr(b2rt_renderer* h, float* rgb, size_t n_floats) {
  if (!h) { set_error("null handle"); return B2RT_ERR_INVALID; }
  const size_t np = (size_t)h->r.width * h->r.height;
  if (!rgb || n_floats < np * 3) { set_error("output buffer too small"); return B2RT_ERR_INVALID; }
  std::vector<float> tmp(np * 4);
  int rc = b2rt_read_rgba32f(h, tmp.data(), tmp.size());
  if (rc) return rc;
  for (size_t i = 0; i < np; ++i) { rgb[3 * i] = tmp[4 * i]; rgb[3 * i + 1] = tmp[4 * i + 1]; rgb[3 * i + 2] = tmp[4 * i + 2]; }
  return B2RT_OK;
}
int b2rt_read_ldr(b2rt_renderer* h, uint32_t* rgba8, size_t n_pixels) {
  int rc = read_common(h, true);
  if (rc) return rc;
  const size_t np = (size_t)h->r.width * h->r.height;
  if (!rgba8 || n_pixels < np) { set_error("output buffer too small"); return B2RT_ERR_INVALID; }
  B2RT_CUDA_OK(cudaMemcpyAsync(rgba8, h->r.ldr, np * 4, cudaMemcpyDeviceToHost, h->r.stream));
  B2RT_CUDA_OK(cudaStreamSynchronize(h->r.stream));
  return B2RT_OK;
}
int b2rt_get_stats(b2rt_renderer* h, b2rt_stats* out) {
  if (!h || !out) { set_error("null argument"); return B2RT_ERR_INVALID; }
  h->r.fill_stats(out);
  return B2RT_OK;
}
int b2rt_accum_device_ptr(b2rt_renderer* h, void** dev_ptr, size_t* n_floats) {
  if (!h || !dev_ptr || !h->r.accum) { set_error("no frame buffer"); return B2RT_ERR_INVALID; }
  *dev_ptr = h->r.accum;
  if (n_floats) *n_floats = (size_t)h->r.width * h->r.height * 4;
  return B2RT_OK;
}
int b2rt_stream_handle(b2rt_renderer* h, void** s) {
  if (!h || !s) { set_error("null argument"); return B2RT_ERR_INVALID; }
  *s = (void*)h->r.stream;
  return B2RT_OK;
}
int b2rt_set_stream(b2rt_renderer* h, void* s) {
  if (!h) { set_error("null handle"); return B2RT_ERR_INVALID; }
  return h->r.set_stream((cudaStream_t)s);
}
int b2rt_set_profiling(b2rt_renderer* h, int collect_counters, int time_kernels) {
  if (!h) { set_error("null handle"); return B2RT_ERR_INVALID; }
  if (h->r.running) { int rc = h->r.wait(); if (rc) return rc; }
  h->r.tracer.collect_stats = collect_counters != 0;
  h->r.tracer.time_kernels = time_kernels != 0;
  return B2RT_OK;
}
void b2rt_destroy(b2rt_renderer* h) {
  if (!h) return;
  h->r.set_device();
  h->r.destroy();
  delete h;
}

}  // extern "C"

// ---- measured FP32 peak ("measure with an FMA-saturating microbenchmark on the box") -----------------
namespace {
__global__ void __launch_bounds__(256) k_fma_peak(float* out, float a, float b, int iters) {
  // 16 independent FFMA chains per thread: enough ILP for one warp to keep its scheduler's FMA pipe fed
  float x[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) x[k] = (float)(threadIdx.x + k);
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int k = 0; k < 16; ++k) x[k] = __fmaf_rn(x[k], a, b);
  }
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < 16; ++k) s += x[k];
  if (s == 123.456f) out[blockIdx.x * blockDim.x + threadIdx.x] = s;   // never true; keeps the chains alive
}
}  // namespace

extern "C" int b2rt_bench_fp32(int32_t device, double* tflops) {
  if (!tflops) { set_error("null argument"); return B2RT_ERR_INVALID; }
  if (!has_device()) { set_error("no CUDA device available (b2rt has no CPU fallback)"); return B2RT_ERR_NO_DEVICE; }
  if (device < 0) B2RT_CUDA_OK(cudaGetDevice(&device));
  B2RT_CUDA_OK(cudaSetDevice(device));
  int sms = 0;
  B2RT_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
  float* d = nullptr;
  const int blocks = sms * 8, iters = 4096;
  B2RT_CUDA_OK(cudaMalloc(&d, (size_t)blocks * 256 * 4));
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  double best = 0;
  for (int rep = 0; rep < 6; ++rep) {   // the first repetitions are the warm-up
    cudaEventRecord(e0);
    k_fma_peak<<<blocks, 256>>>(d, 0.999f, 0.001f, iters);
    cudaEventRecord(e1);
    if (cudaEventSynchronize(e1) != cudaSuccess) break;
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
    const double fl = 2.0 * 16.0 * iters * (double)blocks * 256.0;
    if (ms > 0 && rep >= 2) best = std::max(best, fl / (ms * 1e-3) / 1e12);
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(d);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess || best == 0) { set_error(std::string("b2rt_bench_fp32: ") + cudaGetErrorString(e)); return B2RT_ERR_CUDA; }
  *tflops = best;
  return B2RT_OK;
}

// PathTracer::save_image (src/pathtracer.cpp:577-591) and the HDR counterpart
extern "C" int b2rt_write_png(b2rt_renderer* h, const char* path) {
  if (!h || !path) { set_error("null argument"); return B2RT_ERR_INVALID; }
  std::vector<uint32_t> px((size_t)h->r.width * h->r.height);
  int rc = b2rt_read_ldr(h, px.data(), px.size());
  if (rc) return rc;
  return b2rt_save_png(path, px.data(), h->r.width, h->r.height);
}
extern "C" int b2rt_write_exr(b2rt_renderer* h, const char* path) {
  if (!h || !path) { set_error("null argument"); return B2RT_ERR_INVALID; }
  std::vector<float> px((size_t)h->r.width * h->r.height * 3);
  int rc = b2rt_read_hdr(h, px.data(), px.size());
  if (rc) return rc;
  return b2rt_save_exr(path, px.data(), h->r.width, h->r.height);
}
