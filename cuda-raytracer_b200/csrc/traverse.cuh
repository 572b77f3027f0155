// Device-side traversal + ray scheduler interface (see traverse.cu).
#pragma once
#include <cuda_runtime.h>

#include "b2rt_internal.h"

namespace b2rt {

struct DeviceBVH {
  uint8_t* blob = nullptr;
  TreeletDesc* treelets = nullptr;
  uint32_t n_treelets = 0, n_levels = 0, width = 4, max_treelet_bytes = 0;
  LevelRange levels[MAX_LEVELS];
  uint64_t blob_bytes = 0;
  uint64_t blob_cap = 0, treelet_cap = 0;   // grow-only device allocations (re-upload without cudaMalloc/cudaFree)
};

// Device-resident statistics (all u64): see b2rt_stats
struct TraceCounters {
  unsigned long long node_visits, prim_tests, subtree_visits, pushes;
  unsigned long long staged_bytes;   // subtree bytes moved global -> shared by the TMA bulk copies
  unsigned long long hit_updates;    // 64-bit atomicMin operations issued
};

// Work buffers of the scheduler.  One Tracer serves one stream.
struct Tracer {
  DeviceBVH bvh;
  uint64_t max_rays = 0, pair_cap = 0, chunk_cap = 0;
  uint64_t nt_cap = 0, chunk_alloc = 0;     // grow-only capacities of the per-subtree arrays / chunk list
  uint32_t chunk_rays = 2048;      // rays per work item of a level >= 1 subtree queue (upper bound)
  uint32_t chunk_min = 256;        // ... shrunk down to this when the level has fewer than chunks_per_cta chunks per resident CTA
  uint32_t chunks_per_cta = 2;
  uint32_t count_ctas = 4, scatter_ctas = 6;   // CTAs per SM of k_count_tiled / k_scatter_tiled
  uint32_t chunk0_max = 8192;      // level 0 (one subtree, every ray): chunks grow up to this
  size_t stack_off = 0;            // offset of the traversal stacks in dynamic shared memory
  size_t ring_off = 0;             // offset of the ray ring
  int num_sms = 148, ctas_per_sm = 1;
  size_t smem_bytes = 0;
  // device buffers
  uint32_t* cnt = nullptr;        // [n_treelets] rays queued per subtree
  uint32_t* seg_off = nullptr;    // [n_treelets]
  uint32_t* cursor = nullptr;     // [n_treelets]
  uint32_t* sched_scratch = nullptr;   // k_schedule_level: ticket + per-tile (flag, prefixes)
  uint2* pairs = nullptr;         // [pair_cap] (subtree id, ray id)
  uint32_t* ids_sorted = nullptr; // [pair_cap] ray ids grouped by subtree (levels >= 1)
  uint4* chunks = nullptr;        // [chunk_cap] (subtree, first, count, -)
  uint32_t* ctrl = nullptr;       // [16] pair_count[2], n_chunks, next_chunk, overflow, ...
  TraceCounters* counters = nullptr;   // [2]: level 0 (the root subtree, every ray) and the deeper levels
  bool collect_stats = false;
  uint64_t launches = 0;
  // per-launch CUDA-event timing of k_traverse (the dominant kernel), on the launching stream
  bool time_kernels = false;
  std::vector<cudaEvent_t> ev_pool;
  std::vector<uint8_t> ev_deeper;    // per event pair: 1 = a launch of a level >= 1
  size_t ev_used = 0;
  uint64_t traverse_launches = 0, traverse_launches_l0 = 0;
  // call after the stream is synchronised; resets the pool cursor.  Returns the total; *ms_l0 = the level-0 launches' share
  double harvest_traverse_ms(double* ms_l0 = nullptr);
  // sum of the two counter sets -> *total, the level-0 set -> *l0 (blocking copies)
  int read_counters(TraceCounters* total, TraceCounters* l0);
  int reset_counters(cudaStream_t s);

  int init(const DeviceBVH& b, uint64_t max_rays_, uint32_t pair_factor);
  void release();
  // rays 0..n-1 (a dense list): o = (ox,oy,oz,tmin), d = (dx,dy,dz,tmax); hits must be initialised by the
  // caller to pack(tmax, 0xFFFFFFFF).  n_active_dev: device count of rays.  The arrays must be readable up to the
  // next multiple of 4 entries (TMA tiles are copied in 16-byte units).
  int trace(cudaStream_t s, const float4* ray_o, const float4* ray_d, unsigned long long* hits,
            const uint32_t* n_active_dev, bool any_hit);
  int check_overflow(cudaStream_t s, bool* overflow);  // synchronises the stream

  // ---- distance-sliced tracing (deep subtree graphs) ---------------------------------------------------------
  // The breadth-first scheme has no front-to-back order ACROSS subtrees: a ray is queued at every subtree its whole
  // interval overlaps before any hit is known.  trace_sliced() restores the order at batch granularity: pass p traces
  // only the interval [lo_p, lo_p + first * growth^p] of every ray still without a hit (the last pass takes the rest),
  // so the slab test culls everything behind the slice and rays that hit early never reach the far subtrees.  The
  // result is the same (t, prim) argmin: a hit inside a slice beats everything in the later slices.
  float slice_first = 0.f;         // length of the first slice (0 = slicing off)
  float slice_growth = 4.f;
  int slice_passes = 4;
  float slice_bbox[6] = {0, 0, 0, 0, 0, 0};
  uint64_t slice_cap = 0;
  float4* sl_o[2] = {nullptr, nullptr};
  float4* sl_d[2] = {nullptr, nullptr};
  unsigned long long* sl_h[2] = {nullptr, nullptr};
  uint32_t* sl_map[2] = {nullptr, nullptr};
  float* sl_exit[2] = {nullptr, nullptr};
  uint32_t* sl_n = nullptr;        // [2] device counts of the two lists
  int ensure_slices(uint64_t n);
  // same contract as trace(); n_max = host upper bound of *n_active_dev
  int trace_sliced(cudaStream_t s, const float4* ray_o, const float4* ray_d, unsigned long long* hits,
                   const uint32_t* n_active_dev, uint64_t n_max, bool any_hit);
};

enum { CTRL_PAIRS0 = 0, CTRL_PAIRS1 = 1, CTRL_NEXT0 = 2, CTRL_NEXT1 = 3, CTRL_OVERFLOW = 4, CTRL_NCHUNKS = 5 };

int upload_bvh(const WideBVH& h, DeviceBVH* d);
// device builder (bvh_build_gpu.cu): fills *d directly; *meta gets everything of WideBVH but the blob
// geom_out (optional): device buffer of n_prims * 48 bytes that receives the primitive records in scene order
int build_wide_bvh_device(const b2rt_scene_desc* sc, uint32_t max_leaf, uint32_t width, uint32_t treelet_bytes, cudaStream_t s,
                          DeviceBVH* d, WideBVH* meta, void* geom_out = nullptr);
void free_bvh(DeviceBVH* d);

#define B2RT_CUDA_OK(call)                                                                          \
  do {                                                                                              \
    cudaError_t e__ = (call);                                                                       \
    if (e__ != cudaSuccess) {                                                                       \
      ::b2rt::set_error(std::string(#call) + ": " + cudaGetErrorString(e__));                       \
      return e__ == cudaErrorMemoryAllocation ? B2RT_ERR_OOM : B2RT_ERR_CUDA;                       \
    }                                                                                               \
  } while (0)

}  // namespace b2rt
