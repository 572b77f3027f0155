// Device BVH construction (replaces BVHAccel::BVHAccel + compactTree + compress,
// src/bvh.cpp:339-365, 275-337, 234-273, for scenes where the host build dominates set-up: the 10 M triangle soup
// takes seconds on the host and tens of milliseconds here).
//
// Pipeline, everything on the device; the host reads back a 4-byte count per wide-tree level and per six PLOC rounds:
//   1. k_bounds          primitive boxes -> scene box + summed projected area (block reduction + ordered-int atomics)
//   2. k_morton          63-bit Morton code of the box centre            3. cub::DeviceRadixSort (key, primitive)
//   4. k_ploc_*          binary tree by parallel locally-ordered clustering over the Morton order (Meister & Bittner 2018;
//                        B2RT_GPU_BINARY=lbvh: k_karras, the binary radix tree of Karras 2012, + k_refit)
//   5. k_refit           bottom-up boxes, second arrival at a node proceeds (one atomic counter per node)
//   6. k_collapse        top-down, one launch per wide level: a wide node adopts its binary node's two children and
//                        keeps replacing the largest-area internal child by that child's children until it has W
//                        (same rule as the host builder); binary subtrees of <= max_leaf primitives become leaves
//   7. k_sizes           bottom-up per wide level: bytes / node count / height of every wide subtree
//   8. k_partition       one thread per subtree root, one launch per subtree level: breadth-first packing under the
//                        byte / depth / node budgets with the host builder's rule "a node whose whole subtree fits a
//                        blob of its own is never split", writing finished wide nodes straight into the subtree's slab
//   9. k_compact + k_copy_prims   slabs -> dense blob, primitive records gathered behind each subtree's nodes
// The result is the same DeviceBVH the host builder's upload produces (same node / reference / primitive formats), so
// the traversal kernels and the structural validator do not know which builder ran.  Tree quality is LBVH (spatial
// median), not SAH: good for evenly spread primitives (the soup), measurably worse on meshes inside large boxes.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include "traverse.cuh"

namespace b2rt {
namespace {

__host__ __device__ __forceinline__ uint32_t f2o(float f) {   // order-preserving float -> uint
  uint32_t u;
#ifdef __CUDA_ARCH__
  u = __float_as_uint(f);
#else
  memcpy(&u, &f, 4);
#endif
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__host__ __forceinline__ float o2f(uint32_t o) {
  uint32_t u = (o & 0x80000000u) ? (o & 0x7FFFFFFFu) : ~o;
  float f; memcpy(&f, &u, 4);
  return f;
}

struct SceneBounds { uint32_t lo[3], hi[3]; double projected; };

__device__ __forceinline__ void prim_box(const float4* geom, uint32_t i, uint32_t n_tris, float lo[3], float hi[3]) {
  const float4 a = geom[3 * (size_t)i], b = geom[3 * (size_t)i + 1], c = geom[3 * (size_t)i + 2];
  if (i < n_tris) {
    const float p1[3] = {a.x, a.y, a.z};
    const float p2[3] = {a.x + a.w, a.y + b.x, a.z + b.y};
    const float p3[3] = {a.x + b.z, a.y + b.w, a.z + c.x};
#pragma unroll
    for (int k = 0; k < 3; ++k) { lo[k] = fminf(p1[k], fminf(p2[k], p3[k])); hi[k] = fmaxf(p1[k], fmaxf(p2[k], p3[k])); }
  } else {
    const float p[3] = {a.x, a.y, a.z};
#pragma unroll
    for (int k = 0; k < 3; ++k) { lo[k] = p[k] - a.w; hi[k] = p[k] + a.w; }
  }
}

// scene arrays as the caller holds them (9 floats per triangle, 4 per sphere) -> 48-byte primitive records; the same
// fp32 subtractions as the host path's make_host_scene (e1 = p2 - p1, e2 = p3 - p1, src/static_scene/triangle.cpp:172)
__global__ void __launch_bounds__(256) k_make_prims(const float* __restrict__ tri_verts, const float* __restrict__ spheres,
                                                    uint32_t n_tris, uint32_t n, float4* geom) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (i < n_tris) {
    const float* v = tri_verts + (size_t)i * 9;
    const float v0 = v[0], v1 = v[1], v2 = v[2];
    geom[3 * (size_t)i] = make_float4(v0, v1, v2, v[3] - v0);
    geom[3 * (size_t)i + 1] = make_float4(v[4] - v1, v[5] - v2, v[6] - v0, v[7] - v1);
    geom[3 * (size_t)i + 2] = make_float4(v[8] - v2, __uint_as_float(i), __uint_as_float(0u), 0.f);
  } else {
    const float* sp = spheres + (size_t)(i - n_tris) * 4;
    geom[3 * (size_t)i] = make_float4(sp[0], sp[1], sp[2], sp[3]);
    geom[3 * (size_t)i + 1] = make_float4(0.f, 0.f, 0.f, 0.f);
    geom[3 * (size_t)i + 2] = make_float4(0.f, __uint_as_float(i), __uint_as_float(1u), 0.f);
  }
}

__global__ void __launch_bounds__(256) k_bounds(const float4* __restrict__ geom, uint32_t n, uint32_t n_tris, SceneBounds* out) {
  float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
  double proj = 0;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    float l[3], h[3];
    prim_box(geom, i, n_tris, l, h);
#pragma unroll
    for (int k = 0; k < 3; ++k) { lo[k] = fminf(lo[k], l[k]); hi[k] = fmaxf(hi[k], h[k]); }
    const float4 a = geom[3 * (size_t)i], b = geom[3 * (size_t)i + 1], c = geom[3 * (size_t)i + 2];
    if (i < n_tris) {
      const double e1x = a.w, e1y = b.x, e1z = b.y, e2x = b.z, e2y = b.w, e2z = c.x;
      const double cx = e1y * e2z - e1z * e2y, cy = e1z * e2x - e1x * e2z, cz = e1x * e2y - e1y * e2x;
      proj += 0.25 * sqrt(cx * cx + cy * cy + cz * cz);
    } else {
      proj += 3.14159265358979 * (double)a.w * a.w;
    }
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      lo[k] = fminf(lo[k], __shfl_xor_sync(0xffffffffu, lo[k], d));
      hi[k] = fmaxf(hi[k], __shfl_xor_sync(0xffffffffu, hi[k], d));
    }
    proj += __shfl_xor_sync(0xffffffffu, proj, d);
  }
  if ((threadIdx.x & 31) == 0) {
#pragma unroll
    for (int k = 0; k < 3; ++k) { atomicMin(&out->lo[k], f2o(lo[k])); atomicMax(&out->hi[k], f2o(hi[k])); }
    atomicAdd(&out->projected, proj);
  }
}

__device__ __forceinline__ uint64_t spread21(uint32_t v) {   // 21 bits -> every third bit of 63
  uint64_t x = v & 0x1FFFFFull;
  x = (x | x << 32) & 0x1F00000000FFFFull;
  x = (x | x << 16) & 0x1F0000FF0000FFull;
  x = (x | x << 8) & 0x100F00F00F00F00Full;
  x = (x | x << 4) & 0x10C30C30C30C30C3ull;
  x = (x | x << 2) & 0x1249249249249249ull;
  return x;
}

struct BuildBox { float lo[3], inv[3]; float large2; };   // scene box origin, 2^21 / extent per axis, (large-primitive diagonal)^2

__global__ void __launch_bounds__(256) k_morton(const float4* __restrict__ geom, uint32_t n, uint32_t n_tris, BuildBox bx,
                                                uint64_t* keys, uint32_t* vals) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float l[3], h[3];
  prim_box(geom, i, n_tris, l, h);
  uint32_t q[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const float c = (0.5f * (l[k] + h[k]) - bx.lo[k]) * bx.inv[k];
    q[k] = (uint32_t)fminf(fmaxf(c, 0.f), 2097151.f);
  }
  // Primitives much larger than their neighbours (the walls of a box around a fine mesh) poison a spatial-median tree:
  // sorted by centre they end up deep among the small ones and inflate every ancestor box.  They get the top key bit, so
  // the radix tree's first split separates them into a small subtree of their own next to the root.
  const float dx = h[0] - l[0], dy = h[1] - l[1], dz = h[2] - l[2];
  const uint64_t big = (dx * dx + dy * dy + dz * dz > bx.large2) ? (1ull << 63) : 0ull;
  keys[i] = big | spread21(q[0]) | (spread21(q[1]) << 1) | (spread21(q[2]) << 2);
  vals[i] = i;
}

// ---- binary radix tree.  Unified node ids: internal i in [0, n-1), leaf (sorted position j) = n - 1 + j ----
struct BinTree {
  uint32_t n;
  const uint64_t* keys;
  uint32_t* left; uint32_t* right;        // [n-1] unified ids
  uint32_t* first; uint32_t* last;        // [n-1] sorted range covered
  uint32_t* parent;                       // [2n-1]
  uint32_t* flag;                         // [n-1] refit arrival counters
  float4* box_lo; float4* box_hi;         // [2n-1]
};

__device__ __forceinline__ int delta(const uint64_t* keys, uint32_t n, int i, int j) {
  if (j < 0 || j >= (int)n) return -1;
  const uint64_t a = keys[i], b = keys[j];
  if (a == b) return 64 + __clz((uint32_t)i ^ (uint32_t)j);
  return __clzll((long long)(a ^ b));
}

__global__ void __launch_bounds__(256) k_karras(BinTree T) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int n = (int)T.n;
  if (i >= n - 1) return;
  const int d = (delta(T.keys, n, i, i + 1) - delta(T.keys, n, i, i - 1)) >= 0 ? 1 : -1;
  const int dmin = delta(T.keys, n, i, i - d);
  int lmax = 2;
  while (delta(T.keys, n, i, i + lmax * d) > dmin) lmax <<= 1;
  int l = 0;
  for (int t = lmax >> 1; t >= 1; t >>= 1)
    if (delta(T.keys, n, i, i + (l + t) * d) > dmin) l += t;
  const int j = i + l * d;
  const int dnode = delta(T.keys, n, i, j);
  int s = 0;
  for (int t = (l + 1) >> 1;; t = (t + 1) >> 1) {
    if (delta(T.keys, n, i, i + (s + t) * d) > dnode) s += t;
    if (t <= 1) break;
  }
  const int gamma = i + s * d + min(d, 0);
  const int lo = min(i, j), hi = max(i, j);
  const uint32_t L = (lo == gamma) ? (uint32_t)(n - 1 + gamma) : (uint32_t)gamma;
  const uint32_t R = (hi == gamma + 1) ? (uint32_t)(n - 1 + gamma + 1) : (uint32_t)(gamma + 1);
  T.left[i] = L; T.right[i] = R; T.first[i] = (uint32_t)lo; T.last[i] = (uint32_t)hi;
  T.parent[L] = (uint32_t)i; T.parent[R] = (uint32_t)i;
  if (i == 0) T.parent[0] = 0xFFFFFFFFu;
}

__global__ void __launch_bounds__(256) k_refit(BinTree T, const float4* __restrict__ geom, const uint32_t* __restrict__ sorted,
                                               uint32_t n_tris, float pad) {
  const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= T.n) return;
  float l[3], h[3];
  prim_box(geom, sorted[j], n_tris, l, h);
  float4 lo = make_float4(l[0] - pad, l[1] - pad, l[2] - pad, 0.f), hi = make_float4(h[0] + pad, h[1] + pad, h[2] + pad, 0.f);
  uint32_t id = T.n - 1 + j;
  T.box_lo[id] = lo; T.box_hi[id] = hi;
  if (T.n == 1) return;
  uint32_t p = T.parent[id];
  while (p != 0xFFFFFFFFu) {
    __threadfence();
    if (atomicAdd(&T.flag[p], 1u) == 0u) return;     // first arrival: the sibling's thread finishes the node
    __threadfence();
    const uint32_t a = T.left[p], b = T.right[p];
    const float4 al = __ldcg(&T.box_lo[a]), ah = __ldcg(&T.box_hi[a]), bl = __ldcg(&T.box_lo[b]), bh = __ldcg(&T.box_hi[b]);
    lo = make_float4(fminf(al.x, bl.x), fminf(al.y, bl.y), fminf(al.z, bl.z), 0.f);
    hi = make_float4(fmaxf(ah.x, bh.x), fmaxf(ah.y, bh.y), fmaxf(ah.z, bh.z), 0.f);
    T.box_lo[p] = lo; T.box_hi[p] = hi;
    p = T.parent[p];
  }
}

// ---- PLOC: parallel locally-ordered clustering (Meister & Bittner 2018) -----------------------------------------
// The binary radix tree above splits at the spatial median of the Morton order; its boxes overlap much more than a
// surface-area-heuristic build's (rays on mesh-in-box scenes were 23-45 % slower on it).  PLOC
// builds the binary tree bottom-up instead: the clusters (initially one per primitive, in Morton order) each look for
// the neighbour within +-PLOC_RADIUS positions whose union box has the smallest surface area; mutual nearest
// neighbours merge, the list is compacted, repeat until one cluster is left.  Same unified node ids as the radix tree
// (internal i in [0, n-1), leaves n-1+j; merge k gets id n-2-k, so the root -- the last merge -- is 0); afterwards the
// leaves are renumbered in depth-first order so that every subtree covers a contiguous primitive range (first / last),
// which is what the collapse / partition kernels read.
#ifndef B2RT_PLOC_RADIUS
#define B2RT_PLOC_RADIUS 8
#endif
constexpr int PLOC_RADIUS = B2RT_PLOC_RADIUS;

struct PlocClusters { uint32_t* id; float4* lo; float4* hi; };

__global__ void __launch_bounds__(256) k_ploc_init(BinTree T, const float4* __restrict__ geom, const uint32_t* __restrict__ sorted,
                                                   uint32_t n_tris, float pad, PlocClusters C, uint32_t* size, uint32_t* c_dev) {
  const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j == 0) c_dev[0] = T.n;   // cluster count of round 0 (the later rounds' counts are written by k_ploc_compact)
  if (j >= T.n) return;
  float l[3], h[3];
  prim_box(geom, sorted[j], n_tris, l, h);
  const float4 lo = make_float4(l[0] - pad, l[1] - pad, l[2] - pad, 0.f), hi = make_float4(h[0] + pad, h[1] + pad, h[2] + pad, 0.f);
  const uint32_t id = T.n - 1 + j;
  T.box_lo[id] = lo; T.box_hi[id] = hi;
  T.parent[id] = 0xFFFFFFFFu;
  size[id] = 1u;
  C.id[j] = id; C.lo[j] = lo; C.hi[j] = hi;
}

// The clustering kernels read the round's cluster count from device memory (c_dev) and are launched with a grid sized
// for an UPPER bound of it, so that several rounds can be enqueued back to back without the host reading the count.
__global__ void __launch_bounds__(256) k_ploc_nn(PlocClusters C, const uint32_t* __restrict__ c_dev, uint32_t* nn) {
  const uint32_t c = *c_dev;
  __shared__ float4 s_lo[256 + 2 * PLOC_RADIUS], s_hi[256 + 2 * PLOC_RADIUS];
  const int base = (int)(blockIdx.x * 256) - PLOC_RADIUS;
  for (int k = threadIdx.x; k < 256 + 2 * PLOC_RADIUS; k += 256) {
    const int g = base + k;
    if (g >= 0 && g < (int)c) { s_lo[k] = C.lo[g]; s_hi[k] = C.hi[g]; }
  }
  __syncthreads();
  const uint32_t i = blockIdx.x * 256 + threadIdx.x;
  if (i >= c) return;
  const float4 lo = s_lo[threadIdx.x + PLOC_RADIUS], hi = s_hi[threadIdx.x + PLOC_RADIUS];
  float best = INFINITY; uint32_t bj = 0xFFFFFFFFu;
#pragma unroll 4
  for (int dlt = -PLOC_RADIUS; dlt <= PLOC_RADIUS; ++dlt) {
    const int g = (int)i + dlt;
    if (dlt == 0 || g < 0 || g >= (int)c) continue;
    const float4 a = s_lo[threadIdx.x + PLOC_RADIUS + dlt], b = s_hi[threadIdx.x + PLOC_RADIUS + dlt];
    const float dx = fmaxf(hi.x, b.x) - fminf(lo.x, a.x), dy = fmaxf(hi.y, b.y) - fminf(lo.y, a.y), dz = fmaxf(hi.z, b.z) - fminf(lo.z, a.z);
    const float area = dx * dy + dy * dz + dz * dx;
    if (area < best) { best = area; bj = (uint32_t)g; }     // ascending g: ties go to the smallest index on both sides
  }
  nn[i] = bj;
}

// mutual nearest neighbours merge into a new internal node at the lower position; valid[i] = 1 for positions that stay
__global__ void __launch_bounds__(256) k_ploc_merge(BinTree T, PlocClusters C, const uint32_t* __restrict__ c_dev, uint32_t c_upper,
                                                    const uint32_t* __restrict__ nn, uint32_t* size, uint32_t* merges, uint32_t* valid) {
  const uint32_t c = *c_dev;
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= c) { if (i < c_upper) valid[i] = 0u; return; }   // (the compaction's scan runs over c_upper positions)
  const uint32_t j = nn[i];
  uint32_t keep = 1u;
  if (j != 0xFFFFFFFFu && nn[j] == i) {
    if (i < j) {
      const uint32_t id = T.n - 2u - atomicAdd(merges, 1u);
      const uint32_t a = C.id[i], b = C.id[j];
      const float4 lo = make_float4(fminf(C.lo[i].x, C.lo[j].x), fminf(C.lo[i].y, C.lo[j].y), fminf(C.lo[i].z, C.lo[j].z), 0.f);
      const float4 hi = make_float4(fmaxf(C.hi[i].x, C.hi[j].x), fmaxf(C.hi[i].y, C.hi[j].y), fmaxf(C.hi[i].z, C.hi[j].z), 0.f);
      T.left[id] = a; T.right[id] = b;
      T.parent[a] = id; T.parent[b] = id; T.parent[id] = 0xFFFFFFFFu;
      T.box_lo[id] = lo; T.box_hi[id] = hi;
      size[id] = size[a] + size[b];
      C.id[i] = id; C.lo[i] = lo; C.hi[i] = hi;
    } else {
      keep = 0u;
    }
  }
  valid[i] = keep;
}

__global__ void __launch_bounds__(256) k_ploc_compact(PlocClusters in, const uint32_t* __restrict__ c_dev, const uint32_t* __restrict__ valid,
                                                      const uint32_t* __restrict__ pos, PlocClusters out, uint32_t n,
                                                      const uint32_t* __restrict__ merges, uint32_t* c_next) {
  const uint32_t c = *c_dev;
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0) *c_next = n - *merges;   // every merge of this round is counted: k_ploc_merge has finished
  if (i >= c || !valid[i]) return;
  const uint32_t p = pos[i];
  out.id[p] = in.id[i]; out.lo[p] = in.lo[i]; out.hi[p] = in.hi[i];
}

// The last rounds of the clustering (a few thousand clusters and fewer) in ONE CTA: the same three steps per round,
// separated by __syncthreads instead of kernel boundaries and a host read-back of the cluster count.
constexpr uint32_t PLOC_TAIL = 2048;
__global__ void __launch_bounds__(1024) k_ploc_tail(BinTree T, PlocClusters A, PlocClusters B, uint32_t c, uint32_t* size,
                                                    uint32_t* merges, uint32_t* nn) {
  __shared__ uint32_t s_warp[32];
  __shared__ uint32_t s_total;
  const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  PlocClusters cur = A, nxt = B;
  while (c > 1) {
    for (uint32_t i = tid; i < c; i += 1024) {
      const float4 lo = cur.lo[i], hi = cur.hi[i];
      float best = INFINITY; uint32_t bj = 0xFFFFFFFFu;
      const int g0 = max(0, (int)i - PLOC_RADIUS), g1 = min((int)c - 1, (int)i + PLOC_RADIUS);
      for (int g = g0; g <= g1; ++g) {
        if (g == (int)i) continue;
        const float4 a = cur.lo[g], b = cur.hi[g];
        const float dx = fmaxf(hi.x, b.x) - fminf(lo.x, a.x), dy = fmaxf(hi.y, b.y) - fminf(lo.y, a.y), dz = fmaxf(hi.z, b.z) - fminf(lo.z, a.z);
        const float area = dx * dy + dy * dz + dz * dx;
        if (area < best) { best = area; bj = (uint32_t)g; }
      }
      nn[i] = bj;
    }
    __syncthreads();
    uint32_t keep[2] = {0u, 0u};
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const uint32_t i = tid + k * 1024u;
      if (i >= c) continue;
      const uint32_t j = nn[i];
      keep[k] = 1u;
      if (j != 0xFFFFFFFFu && nn[j] == i) {
        if (i < j) {
          const uint32_t id = T.n - 2u - atomicAdd(merges, 1u);
          const uint32_t a = cur.id[i], b = cur.id[j];
          const float4 la = cur.lo[i], lb = cur.lo[j], ha = cur.hi[i], hb = cur.hi[j];
          const float4 lo = make_float4(fminf(la.x, lb.x), fminf(la.y, lb.y), fminf(la.z, lb.z), 0.f);
          const float4 hi = make_float4(fmaxf(ha.x, hb.x), fmaxf(ha.y, hb.y), fmaxf(ha.z, hb.z), 0.f);
          T.left[id] = a; T.right[id] = b;
          T.parent[a] = id; T.parent[b] = id; T.parent[id] = 0xFFFFFFFFu;
          T.box_lo[id] = lo; T.box_hi[id] = hi;
          size[id] = size[a] + size[b];
          cur.id[i] = id; cur.lo[i] = lo; cur.hi[i] = hi;
        } else {
          keep[k] = 0u;
        }
      }
    }
    __syncthreads();   // every merged cluster is written before the compaction reads it
    uint32_t base = 0;
#pragma unroll
    for (int k = 0; k < 2; ++k) {   // exclusive scan of the keep flags in index order: items tid, then items tid + 1024
      const uint32_t m = __ballot_sync(0xffffffffu, keep[k] != 0u);
      if (lane == 0) s_warp[warp] = __popc(m);
      __syncthreads();
      if (warp == 0) {
        const uint32_t v = s_warp[lane];
        uint32_t incl = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= (uint32_t)d) incl += t; }
        s_warp[lane] = incl - v;
        if (lane == 31) s_total = incl;
      }
      __syncthreads();
      const uint32_t i = tid + k * 1024u;
      if (keep[k]) {
        const uint32_t p = base + s_warp[warp] + __popc(m & ((1u << lane) - 1u));
        nxt.id[p] = cur.id[i]; nxt.lo[p] = cur.lo[i]; nxt.hi[p] = cur.hi[i];
      }
      base += s_total;
      __syncthreads();
    }
    c = base;
    const PlocClusters t = cur; cur = nxt; nxt = t;
  }
}

// depth-first position of every node: the number of leaves left of it = sum, over the ancestors it hangs under on the
// right, of the left sibling's leaf count
__global__ void __launch_bounds__(256) k_ploc_first(BinTree T, const uint32_t* __restrict__ size, uint32_t* first_of) {
  const uint32_t id = blockIdx.x * blockDim.x + threadIdx.x;
  if (id >= 2u * T.n - 1u) return;
  uint32_t off = 0;
  for (uint32_t c = id, p = T.parent[c]; p != 0xFFFFFFFFu; c = p, p = T.parent[p])
    if (T.right[p] == c) off += size[T.left[p]];
  first_of[id] = off;
}

// final arrays: leaves renumbered by depth-first position (unified id n-1+pos), ranges of the internal nodes
__global__ void __launch_bounds__(256) k_ploc_finish(BinTree T, const uint32_t* __restrict__ size, const uint32_t* __restrict__ first_of,
                                                     const uint32_t* __restrict__ sorted_in, uint32_t* sorted_out,
                                                     float4* leaf_lo, float4* leaf_hi) {
  const uint32_t id = blockIdx.x * blockDim.x + threadIdx.x;
  if (id >= 2u * T.n - 1u) return;
  const uint32_t n1 = T.n - 1u;
  if (id < n1) {
    const uint32_t a = T.left[id], b = T.right[id];
    T.first[id] = first_of[id]; T.last[id] = first_of[id] + size[id] - 1u;
    if (a >= n1) T.left[id] = n1 + first_of[a];
    if (b >= n1) T.right[id] = n1 + first_of[b];
  } else {
    const uint32_t p = first_of[id];
    sorted_out[p] = sorted_in[id - n1];
    leaf_lo[p] = T.box_lo[id]; leaf_hi[p] = T.box_hi[id];
  }
}
__global__ void __launch_bounds__(256) k_ploc_leaf_boxes(BinTree T, const float4* __restrict__ leaf_lo, const float4* __restrict__ leaf_hi) {
  const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= T.n) return;
  T.box_lo[T.n - 1u + j] = leaf_lo[j]; T.box_hi[T.n - 1u + j] = leaf_hi[j];
}

// ---- wide collapse ----
struct WideTmp {
  uint8_t* nodes;          // [cap] temporary wide nodes in the final node layout; INTERNAL payload = global wide index,
                           // LEAF payload = unified binary id of the leaf subtree
  uint32_t* own_prims;     // [cap] primitives in this node's leaf children
  uint32_t* sub_bytes;     // [cap] bytes of the whole wide subtree (saturating)
  uint32_t* sub_nodes;     // [cap]
  uint32_t* height;        // [cap]
  uint32_t* count;         // wide nodes allocated so far
};

__device__ __forceinline__ bool bin_is_leaf(const BinTree& T, uint32_t id, uint32_t max_leaf) {
  return id >= T.n - 1 || (T.last[id] - T.first[id] + 1u) <= max_leaf;
}
__device__ __forceinline__ float box_area(float4 lo, float4 hi) {
  const float dx = hi.x - lo.x, dy = hi.y - lo.y, dz = hi.z - lo.z;
  return 2.f * (dx * dy + dy * dz + dz * dx);
}

template <int W>
__global__ void __launch_bounds__(128) k_collapse(BinTree T, WideTmp Wt, const uint32_t* __restrict__ cur_q, uint32_t level_first,
                                                  uint32_t level_count, uint32_t next_first, uint32_t* next_q, uint32_t cap,
                                                  uint32_t max_leaf, uint32_t* overflow) {
  const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= level_count) return;
  constexpr uint32_t NB = 32u * W + (uint32_t)B2RT_NODE_PAD;   // = node_bytes(W)
  const uint32_t bin = cur_q[k];
  uint32_t kids[W]; float area[W]; int nk = 0;
  if (bin_is_leaf(T, bin, max_leaf)) { kids[nk++] = bin; }
  else { kids[nk++] = T.left[bin]; kids[nk++] = T.right[bin]; }
  for (int q = 0; q < nk; ++q)
    area[q] = bin_is_leaf(T, kids[q], max_leaf) ? -1.f : box_area(T.box_lo[kids[q]], T.box_hi[kids[q]]);
  while (nk < W) {
    int best = -1; float ba = -1.f;
    for (int q = 0; q < nk; ++q) if (area[q] > ba) { ba = area[q]; best = q; }
    if (best < 0) break;
    const uint32_t c = kids[best];
    const uint32_t l = T.left[c], r = T.right[c];
    kids[best] = l; kids[nk] = r;
    area[best] = bin_is_leaf(T, l, max_leaf) ? -1.f : box_area(T.box_lo[l], T.box_hi[l]);
    area[nk] = bin_is_leaf(T, r, max_leaf) ? -1.f : box_area(T.box_lo[r], T.box_hi[r]);
    ++nk;
  }
  float* f = reinterpret_cast<float*>(Wt.nodes + (size_t)(level_first + k) * NB);
  uint32_t* refs = reinterpret_cast<uint32_t*>(f + 6 * W);
  uint32_t own = 0;
  for (int q = 0; q < W; ++q) {
    if (q < nk) {
      const uint32_t c = kids[q];
      const float4 lo = T.box_lo[c], hi = T.box_hi[c];
      f[0 * W + q] = lo.x; f[1 * W + q] = lo.y; f[2 * W + q] = lo.z;
      f[3 * W + q] = hi.x; f[4 * W + q] = hi.y; f[5 * W + q] = hi.z;
      if (area[q] < 0.f) {
        refs[q] = (REF_LEAF << 30) | c;
        own += c >= T.n - 1 ? 1u : (T.last[c] - T.first[c] + 1u);
      } else {
        const uint32_t idx = atomicAdd(Wt.count, 1u);
        if (idx >= cap) { *overflow = 1; refs[q] = REF_EMPTY_WORD; continue; }
        next_q[idx - next_first] = c;
        refs[q] = (REF_INTERNAL << 30) | idx;
      }
    } else {
      f[0 * W + q] = INFINITY; f[1 * W + q] = INFINITY; f[2 * W + q] = INFINITY;
      f[3 * W + q] = -INFINITY; f[4 * W + q] = -INFINITY; f[5 * W + q] = -INFINITY;
      refs[q] = REF_EMPTY_WORD;
    }
  }
  Wt.own_prims[level_first + k] = own;
}

template <int W>
__global__ void __launch_bounds__(256) k_sizes(WideTmp Wt, uint32_t level_first, uint32_t level_count) {
  const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= level_count) return;
  constexpr uint32_t NB = 32u * W + (uint32_t)B2RT_NODE_PAD;   // = node_bytes(W)
  const uint32_t idx = level_first + k;
  const uint32_t* refs = reinterpret_cast<const uint32_t*>(Wt.nodes + (size_t)idx * NB + 24 * W);
  uint64_t bytes = NB + (uint64_t)Wt.own_prims[idx] * PRIM_BYTES;
  uint64_t nodes = 1; uint32_t h = 1;
  for (int q = 0; q < W; ++q) {
    const uint32_t r = refs[q];
    if (r != REF_EMPTY_WORD && (r >> 30) == REF_INTERNAL) {
      const uint32_t c = r & 0x3FFFFFFFu;
      bytes += Wt.sub_bytes[c]; nodes += Wt.sub_nodes[c]; h = max(h, Wt.height[c] + 1u);
    }
  }
  Wt.sub_bytes[idx] = (uint32_t)min(bytes, (uint64_t)0xFFFFFFFFu);
  Wt.sub_nodes[idx] = (uint32_t)min(nodes, (uint64_t)0xFFFFFFFFu);
  Wt.height[idx] = h;
}

// ---- partition into subtrees + serialisation ----
struct PartParams {
  uint8_t* slabs; uint32_t stride;           // one slab of `stride` bytes per subtree id
  TreeletDesc* descs;
  const uint32_t* roots; uint32_t n_roots; uint32_t first_id;   // this level: wide index of each subtree root
  uint32_t* next_roots; uint32_t* next_count; uint32_t next_first_id; uint32_t cap_subtrees;
  uint2* prim_dest;                          // [n] sorted position -> (subtree id, local primitive index)
  uint32_t budget, depth_limit, node_limit;
  uint32_t* overflow;
};

// One WARP per subtree root.  The packing itself is sequential (every decision moves the running byte / node counts),
// but what made the one-thread version slow was three dependent global-memory latencies per node (queue entry -> child
// references -> the children's subtree sizes): 286 us per launch on the cfg3 stand-in whatever the number of roots.
// Here the breadth-first queue lives in shared memory and a batch of 32 / W queue nodes is loaded at once, one lane per
// (node, child); the decisions then run over the batch in the same order as before with the lanes' values fetched by
// shuffles (every lane keeps the same running state), the primitives of a leaf child are placed by all lanes, and the
// batch's exits take ONE atomic.  Same decisions in the same order, so the slabs are byte-identical to the old kernel's.
constexpr int PART_WARPS = 2;
template <int W> constexpr uint32_t PART_QUEUE_V = max_treelet_nodes(W);
template <int W>
__global__ void __launch_bounds__(PART_WARPS * 32) k_partition(BinTree T, WideTmp Wt, PartParams P) {
  constexpr uint32_t NB = 32u * W + (uint32_t)B2RT_NODE_PAD;   // = node_bytes(W)
  constexpr uint32_t PAD = 28 * W;            // byte offset of a node's unused tail
  constexpr uint32_t NPB = 32 / W;            // queue nodes per batch
  constexpr uint32_t ROWS16 = 24 * W / 16;    // 16-byte words of a node's box rows
  constexpr uint32_t FULL = 0xffffffffu;
  __shared__ uint2 s_queue[PART_WARPS][PART_QUEUE_V<W>];   // (wide index, depth | whole << 8)
  __shared__ uint2 s_exit[PART_WARPS][32];                      // this batch's exits: (slab byte offset of the reference, wide index)
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t k = blockIdx.x * PART_WARPS + warp;
  if (k >= P.n_roots) return;
  const uint32_t tid = P.first_id + k;
  uint8_t* slab = P.slabs + (size_t)tid * P.stride;
  uint2* Q = s_queue[warp];
  uint2* X = s_exit[warp];
  const uint32_t root = P.roots[k];
  uint32_t n_nodes = 1, committed = 1, n_prims = 0;
  uint64_t used = NB + (uint64_t)Wt.own_prims[root] * PRIM_BYTES;
  if (lane == 0) Q[0] = make_uint2(root, 1u);
  __syncwarp();
  for (uint32_t q0 = 0, nb = 0; q0 < n_nodes; q0 += nb) {
    nb = min(NPB, n_nodes - q0);   // the queue grows while the batch is decided; the next batch starts where this one ends
    // ---- load phase: lane (j, c) reads child c of queue node q0 + j and what the decision needs to know about it
    const uint32_t j = lane / W, c = lane % W;
    const bool valid = j < nb;
    const uint2 qe = valid ? Q[q0 + j] : make_uint2(0u, 0u);
    uint32_t r = REF_EMPTY_WORD, m0 = 0, m1 = 0, m2 = 0, m3 = 0;
    if (valid) r = reinterpret_cast<const uint32_t*>(Wt.nodes + (size_t)qe.x * NB + 24 * W)[c];
    if (r != REF_EMPTY_WORD) {
      const uint32_t pay = r & 0x3FFFFFFFu;
      if ((r >> 30) == REF_LEAF) {
        m0 = pay >= T.n - 1 ? pay - (T.n - 1) : T.first[pay];                // first sorted position
        m1 = pay >= T.n - 1 ? 1u : T.last[pay] - m0 + 1u;                    // primitive count
      } else {
        m0 = Wt.sub_bytes[pay]; m1 = Wt.sub_nodes[pay]; m2 = Wt.height[pay]; m3 = Wt.own_prims[pay];
      }
    }
    // box rows of the batch's nodes, tails cleared
    for (uint32_t i = lane; i < nb * ROWS16; i += 32) {
      const uint32_t jj = i / ROWS16, v = i % ROWS16;
      reinterpret_cast<uint4*>(slab + (size_t)(q0 + jj) * NB)[v] = reinterpret_cast<const uint4*>(Wt.nodes + (size_t)Q[q0 + jj].x * NB)[v];
    }
    if (lane < nb) { uint32_t* tail = reinterpret_cast<uint32_t*>(slab + (size_t)(q0 + lane) * NB + PAD); tail[0] = 0; tail[1] = 0; }
    // ---- decisions, in queue / child order
    uint32_t n_ex = 0;
    for (uint32_t jj = 0; jj < nb; ++jj) {
      const uint32_t dw = __shfl_sync(FULL, qe.y, jj * W);
      const uint32_t depth = dw & 0xFFu; const bool whole = (dw >> 8) != 0;
      const uint32_t ref_off = (q0 + jj) * NB + 24 * W;   // slab byte offset of the node's child references
#pragma unroll
      for (uint32_t cc = 0; cc < (uint32_t)W; ++cc) {
        const uint32_t src = jj * W + cc;
        const uint32_t rr = __shfl_sync(FULL, r, src);
        const uint32_t a0 = __shfl_sync(FULL, m0, src), a1 = __shfl_sync(FULL, m1, src);
        const uint32_t a2 = __shfl_sync(FULL, m2, src), a3 = __shfl_sync(FULL, m3, src);
        uint32_t* dref = reinterpret_cast<uint32_t*>(slab + ref_off) + cc;
        if (rr == REF_EMPTY_WORD) { if (lane == 0) *dref = rr; continue; }
        const uint32_t pay = rr & 0x3FFFFFFFu;
        if ((rr >> 30) == REF_LEAF) {
          if (lane == 0) *dref = (REF_LEAF << 30) | ((a1 - 1u) << 24) | n_prims;
          for (uint32_t p = lane; p < a1; p += 32) P.prim_dest[a0 + p] = make_uint2(tid, n_prims + p);
          n_prims += a1;
          continue;
        }
        // internal child: whole subtree / partial / exit (host builder's rules, breadth-first instead of by area)
        const uint32_t sb = a0, sn = a1, sh = a2;
        bool take = whole, take_whole = whole;
        if (!take) {
          if (used + sb <= P.budget && depth + sh <= P.depth_limit && (uint64_t)committed + sn <= P.node_limit) {
            take = take_whole = true; used += sb; committed += sn;
          } else {
            const bool fits_alone = sb <= P.budget && sh <= P.depth_limit && sn <= P.node_limit;
            const uint64_t own = NB + (uint64_t)a3 * PRIM_BYTES;
            if (!fits_alone && used + own <= P.budget && depth + 1 <= P.depth_limit && committed < P.node_limit) {
              take = true; used += own; committed += 1;
            }
          }
        }
        if (take) {
          if (lane == 0) { Q[n_nodes] = make_uint2(pay, (depth + 1u) | (take_whole ? 0x100u : 0u)); *dref = (REF_INTERNAL << 30) | n_nodes; }
          ++n_nodes;
        } else {
          if (lane == 0) X[n_ex] = make_uint2(ref_off + cc * 4u, pay);
          ++n_ex;
        }
      }
    }
    __syncwarp();
    // ---- the batch's exits become subtree roots of the next level: one reservation
    if (n_ex) {
      uint32_t e0 = 0;
      if (lane == 0) e0 = atomicAdd(P.next_count, n_ex);
      e0 = __shfl_sync(FULL, e0, 0);
      if (lane < n_ex) {
        const uint2 x = X[lane];
        uint32_t* dref = reinterpret_cast<uint32_t*>(slab + x.x);
        if (P.next_first_id + e0 + lane >= P.cap_subtrees) { *P.overflow = 2; *dref = REF_EMPTY_WORD; }
        else { P.next_roots[e0 + lane] = x.y; *dref = (REF_EXIT << 30) | (P.next_first_id + e0 + lane); }
      }
      __syncwarp();
    }
  }
  if (lane == 0) {
    TreeletDesc d;
    d.offset16 = 0;   // assigned by the compaction
    d.bytes = (uint32_t)(((uint64_t)n_nodes * NB + (uint64_t)n_prims * PRIM_BYTES + 15u) & ~15ull);
    d.n_nodes = n_nodes; d.n_prims = n_prims;
    P.descs[tid] = d;
  }
}

// slabs -> dense blob (one CTA per subtree, node part only; primitives are gathered by k_copy_prims)
__global__ void __launch_bounds__(128) k_compact(const uint8_t* __restrict__ slabs, uint32_t stride, const TreeletDesc* __restrict__ descs,
                                                 uint32_t node_bytes_, uint8_t* blob) {
  const uint32_t t = blockIdx.x;
  const TreeletDesc d = descs[t];
  const uint4* src = reinterpret_cast<const uint4*>(slabs + (size_t)t * stride);
  uint4* dst = reinterpret_cast<uint4*>(blob + (size_t)d.offset16 * 16);
  const uint32_t n16 = d.n_nodes * (node_bytes_ / 16);
  for (uint32_t i = threadIdx.x; i < n16; i += blockDim.x) dst[i] = src[i];
}

__global__ void __launch_bounds__(256) k_copy_prims(const float4* __restrict__ geom, const uint32_t* __restrict__ sorted,
                                                    const uint2* __restrict__ prim_dest, const TreeletDesc* __restrict__ descs,
                                                    uint32_t n, uint32_t node_bytes_, uint8_t* blob) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint2 pd = prim_dest[i];
  const TreeletDesc d = descs[pd.x];
  float4* dst = reinterpret_cast<float4*>(blob + (size_t)d.offset16 * 16 + (size_t)d.n_nodes * node_bytes_ + (size_t)pd.y * PRIM_BYTES);
  const float4* src = geom + 3 * (size_t)sorted[i];
  dst[0] = src[0]; dst[1] = src[1]; dst[2] = src[2];
}

// Scratch buffers come from the device's stream-ordered pool (cudaMallocAsync) with the release threshold lifted, so a
// rebuild reuses the previous build's memory instead of paying cudaMalloc / cudaFree (tens of milliseconds for GBs).
struct DevBuf {
  cudaStream_t s;
  std::vector<void*> ptrs;
  explicit DevBuf(cudaStream_t s_) : s(s_) {
    int dev = 0; cudaGetDevice(&dev);
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
      uint64_t keep = ~0ull;
      cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
  }
  ~DevBuf() { for (void* p : ptrs) cudaFreeAsync(p, s); }
  template <typename T>
  cudaError_t alloc(T** p, size_t count) {
    cudaError_t e = cudaMallocAsync(p, std::max<size_t>(count, 1) * sizeof(T), s);
    if (e == cudaSuccess) ptrs.push_back(*p);
    return e;
  }
};

}  // namespace

template <int W>
static int build_device_w(const b2rt_scene_desc& sc, uint32_t max_leaf, uint32_t treelet_bytes, cudaStream_t s, DeviceBVH* dev, WideBVH* meta,
                          float4* geom_out) {
  auto t0 = std::chrono::steady_clock::now();
  const bool verbose = getenv("B2RT_VERBOSE") != nullptr;
  auto lap = [&](const char* what) {
    if (!verbose) return;
    cudaStreamSynchronize(s);
    fprintf(stderr, "b2rt: gpu build %-10s %8.3f ms\n", what, std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count());
  };
  constexpr uint32_t NB = 32u * W + (uint32_t)B2RT_NODE_PAD;   // = node_bytes(W)
  const uint32_t n = sc.n_tris + sc.n_spheres;
  const uint32_t depth_limit = stack_entries(W) / (W - 1);
  const uint32_t node_limit = max_treelet_nodes(W);
  DevBuf buf(s);
  float4* geom = nullptr;
  { float *d_tv = nullptr, *d_sp = nullptr;
    if (geom_out) geom = geom_out;   // the caller keeps the 48-byte primitive records (the renderer shades from them)
    else B2RT_CUDA_OK(buf.alloc(&geom, (size_t)n * 3));
    B2RT_CUDA_OK(buf.alloc(&d_tv, (size_t)sc.n_tris * 9));
    B2RT_CUDA_OK(buf.alloc(&d_sp, (size_t)sc.n_spheres * 4));
    if (sc.n_tris) B2RT_CUDA_OK(cudaMemcpyAsync(d_tv, sc.tri_verts, (size_t)sc.n_tris * 36, cudaMemcpyHostToDevice, s));
    if (sc.n_spheres) B2RT_CUDA_OK(cudaMemcpyAsync(d_sp, sc.spheres, (size_t)sc.n_spheres * 16, cudaMemcpyHostToDevice, s));
    k_make_prims<<<(n + 255) / 256, 256, 0, s>>>(d_tv, d_sp, sc.n_tris, n, geom); }
  // 1. scene bounds
  SceneBounds* d_sb = nullptr;
  B2RT_CUDA_OK(buf.alloc(&d_sb, 1));
  SceneBounds sb;
  for (int k = 0; k < 3; ++k) { sb.lo[k] = 0xFFFFFFFFu; sb.hi[k] = 0u; }
  sb.projected = 0;
  B2RT_CUDA_OK(cudaMemcpyAsync(d_sb, &sb, sizeof sb, cudaMemcpyHostToDevice, s));
  int dev_id = 0, n_sms = 148;
  cudaGetDevice(&dev_id);
  cudaDeviceGetAttribute(&n_sms, cudaDevAttrMultiProcessorCount, dev_id);
  k_bounds<<<n_sms * 8, 256, 0, s>>>(geom, n, sc.n_tris, d_sb);
  B2RT_CUDA_OK(cudaMemcpyAsync(&sb, d_sb, sizeof sb, cudaMemcpyDeviceToHost, s));
  B2RT_CUDA_OK(cudaStreamSynchronize(s));
  float lo[3], hi[3];
  for (int k = 0; k < 3; ++k) { lo[k] = o2f(sb.lo[k]); hi[k] = o2f(sb.hi[k]); }
  float diag = 0.f, maxabs = 0.f;
  { const float ex = hi[0] - lo[0], ey = hi[1] - lo[1], ez = hi[2] - lo[2];
    diag = std::sqrt(ex * ex + ey * ey + ez * ez);
    for (int k = 0; k < 3; ++k) maxabs = std::max(maxabs, std::max(std::fabs(lo[k]), std::fabs(hi[k]))); }
  const float pad = std::max(1e-30f, 1e-5f * std::max(diag, maxabs));   // same padding rule as the host builder
  for (int k = 0; k < 3; ++k) { meta->bbox[k] = lo[k]; meta->bbox[3 + k] = hi[k]; }
  meta->mean_free_path = 0.f;
  if (sb.projected > 0) meta->mean_free_path = (float)((double)(hi[0] - lo[0]) * (hi[1] - lo[1]) * (hi[2] - lo[2]) / sb.projected);
  lap("bounds");
  // 2./3. Morton codes, sort
  uint64_t *keys_a = nullptr, *keys_b = nullptr; uint32_t *vals_a = nullptr, *vals_b = nullptr;
  B2RT_CUDA_OK(buf.alloc(&keys_a, n)); B2RT_CUDA_OK(buf.alloc(&keys_b, n));
  B2RT_CUDA_OK(buf.alloc(&vals_a, n)); B2RT_CUDA_OK(buf.alloc(&vals_b, n));
  BuildBox bx;
  for (int k = 0; k < 3; ++k) { bx.lo[k] = lo[k]; const float e = hi[k] - lo[k]; bx.inv[k] = e > 0.f ? 2097152.f / e : 0.f; }
  { float f = 0.125f; if (const char* e = getenv("B2RT_LARGE_PRIM")) f = (float)atof(e);   // fraction of the scene diagonal; 0 = off
    bx.large2 = f > 0.f ? (f * diag) * (f * diag) : INFINITY; }
  k_morton<<<(n + 255) / 256, 256, 0, s>>>(geom, n, sc.n_tris, bx, keys_a, vals_a);
  size_t temp_bytes = 0;
  B2RT_CUDA_OK(cub::DeviceRadixSort::SortPairs(nullptr, temp_bytes, keys_a, keys_b, vals_a, vals_b, (int)n, 0, 64, s));
  uint8_t* temp = nullptr;
  B2RT_CUDA_OK(buf.alloc(&temp, temp_bytes));
  B2RT_CUDA_OK(cub::DeviceRadixSort::SortPairs(temp, temp_bytes, keys_a, keys_b, vals_a, vals_b, (int)n, 0, 64, s));
  const uint64_t* keys = keys_b; const uint32_t* sorted = vals_b;
  lap("sort");
  // 4./5. radix tree + boxes
  BinTree T;
  T.n = n; T.keys = keys;
  B2RT_CUDA_OK(buf.alloc(&T.left, n)); B2RT_CUDA_OK(buf.alloc(&T.right, n));
  B2RT_CUDA_OK(buf.alloc(&T.first, n)); B2RT_CUDA_OK(buf.alloc(&T.last, n));
  B2RT_CUDA_OK(buf.alloc(&T.parent, (size_t)2 * n)); B2RT_CUDA_OK(buf.alloc(&T.flag, n));
  B2RT_CUDA_OK(buf.alloc(&T.box_lo, (size_t)2 * n)); B2RT_CUDA_OK(buf.alloc(&T.box_hi, (size_t)2 * n));
  B2RT_CUDA_OK(cudaMemsetAsync(T.flag, 0, (size_t)n * 4, s));
  // binary tree over the Morton order: PLOC (default; B2RT_GPU_BINARY=lbvh selects the radix tree)
  bool use_ploc = n > 2;
  if (const char* e = getenv("B2RT_GPU_BINARY")) use_ploc = use_ploc && strcmp(e, "lbvh") != 0;
  if (!use_ploc) {
    if (n > 1) k_karras<<<(n + 255) / 256, 256, 0, s>>>(T);
    k_refit<<<(n + 255) / 256, 256, 0, s>>>(T, geom, sorted, sc.n_tris, pad);
    lap("radix tree");
  } else {
    PlocClusters C[2];
    uint32_t *nn = nullptr, *valid = nullptr, *pos = nullptr, *size = nullptr, *first_of = nullptr, *merges = nullptr, *sorted2 = nullptr;
    float4 *leaf_lo = nullptr, *leaf_hi = nullptr;
    for (int k = 0; k < 2; ++k) { B2RT_CUDA_OK(buf.alloc(&C[k].id, n)); B2RT_CUDA_OK(buf.alloc(&C[k].lo, n)); B2RT_CUDA_OK(buf.alloc(&C[k].hi, n)); }
    B2RT_CUDA_OK(buf.alloc(&nn, n)); B2RT_CUDA_OK(buf.alloc(&valid, n)); B2RT_CUDA_OK(buf.alloc(&pos, n));
    B2RT_CUDA_OK(buf.alloc(&size, (size_t)2 * n)); B2RT_CUDA_OK(buf.alloc(&first_of, (size_t)2 * n)); B2RT_CUDA_OK(buf.alloc(&merges, 1));
    B2RT_CUDA_OK(buf.alloc(&sorted2, n)); B2RT_CUDA_OK(buf.alloc(&leaf_lo, n)); B2RT_CUDA_OK(buf.alloc(&leaf_hi, n));
    size_t scan_bytes = 0;
    B2RT_CUDA_OK(cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, valid, pos, (int)n, s));
    uint8_t* scan_tmp = nullptr;
    B2RT_CUDA_OK(buf.alloc(&scan_tmp, scan_bytes));
    B2RT_CUDA_OK(cudaMemsetAsync(merges, 0, 4, s));
    constexpr int ROUNDS_PER_SYNC = 6, MAX_ROUNDS = 8192;
    uint32_t* c_dev = nullptr;
    B2RT_CUDA_OK(buf.alloc(&c_dev, (size_t)MAX_ROUNDS + ROUNDS_PER_SYNC + 1));
    k_ploc_init<<<(n + 255) / 256, 256, 0, s>>>(T, geom, sorted, sc.n_tris, pad, C[0], size, c_dev);
    uint32_t c = n, done = 0;   // c: the exact cluster count at the last read-back = an upper bound for the rounds after it
    int cur = 0, iters = 0, syncs = 0;
    while (c > 1) {
      if (c <= PLOC_TAIL) {   // the remaining rounds in one CTA, no host round trips
        k_ploc_tail<<<1, 1024, 0, s>>>(T, C[cur], C[cur ^ 1], c, size, merges, nn);
        ++iters;
        break;
      }
      // a batch of rounds between two read-backs of the count (a read-back costs more than a round of a mid-sized scene)
      const uint32_t g = (c + 255) / 256;
      for (int r = 0; r < ROUNDS_PER_SYNC; ++r, ++iters) {
        k_ploc_nn<<<g, 256, 0, s>>>(C[cur], c_dev + iters, nn);
        k_ploc_merge<<<g, 256, 0, s>>>(T, C[cur], c_dev + iters, c, nn, size, merges, valid);
        B2RT_CUDA_OK(cub::DeviceScan::ExclusiveSum(scan_tmp, scan_bytes, valid, pos, (int)c, s));
        k_ploc_compact<<<g, 256, 0, s>>>(C[cur], c_dev + iters, valid, pos, C[cur ^ 1], n, merges, c_dev + iters + 1);
        cur ^= 1;
      }
      B2RT_CUDA_OK(cudaMemcpyAsync(&done, merges, 4, cudaMemcpyDeviceToHost, s));
      B2RT_CUDA_OK(cudaStreamSynchronize(s));
      ++syncs;
      if (n - done >= c) { set_error("gpu bvh build: PLOC made no progress"); return B2RT_ERR_INVALID; }
      c = n - done;
      if (iters > MAX_ROUNDS) { set_error("gpu bvh build: PLOC did not converge"); return B2RT_ERR_INVALID; }
    }
    const uint32_t n_all = 2 * n - 1;
    k_ploc_first<<<(n_all + 255) / 256, 256, 0, s>>>(T, size, first_of);
    k_ploc_finish<<<(n_all + 255) / 256, 256, 0, s>>>(T, size, first_of, sorted, sorted2, leaf_lo, leaf_hi);
    k_ploc_leaf_boxes<<<(n + 255) / 256, 256, 0, s>>>(T, leaf_lo, leaf_hi);
    sorted = sorted2;
    if (verbose) fprintf(stderr, "b2rt: gpu build PLOC: %d rounds, %d host read-backs of the cluster count (the last round finishes in one CTA)\n", iters, syncs);
    lap("ploc tree");
  }
  // 6. wide collapse, one launch per wide level
  const uint32_t cap = n + 1;
  WideTmp Wt;
  B2RT_CUDA_OK(buf.alloc(&Wt.nodes, (size_t)cap * NB));
  B2RT_CUDA_OK(buf.alloc(&Wt.own_prims, cap)); B2RT_CUDA_OK(buf.alloc(&Wt.sub_bytes, cap));
  B2RT_CUDA_OK(buf.alloc(&Wt.sub_nodes, cap)); B2RT_CUDA_OK(buf.alloc(&Wt.height, cap));
  uint32_t* ctr = nullptr;    // [0] wide count, [1] overflow, [2] next subtree count
  B2RT_CUDA_OK(buf.alloc(&ctr, 4));
  Wt.count = ctr;
  uint32_t *q_a = nullptr, *q_b = nullptr;
  B2RT_CUDA_OK(buf.alloc(&q_a, cap)); B2RT_CUDA_OK(buf.alloc(&q_b, cap));
  const uint32_t root_bin = n == 1 ? 0u /* unified id of the only leaf */ : 0u;
  { uint32_t init[4] = {1u, 0u, 0u, 0u};
    B2RT_CUDA_OK(cudaMemcpyAsync(ctr, init, 16, cudaMemcpyHostToDevice, s));
    B2RT_CUDA_OK(cudaMemcpyAsync(q_a, &root_bin, 4, cudaMemcpyHostToDevice, s)); }
  std::vector<std::pair<uint32_t, uint32_t>> wide_levels;   // (first, count)
  uint32_t level_first = 0, level_count = 1;
  while (level_count) {
    wide_levels.push_back({level_first, level_count});
    const uint32_t next_first = level_first + level_count;
    k_collapse<W><<<(level_count + 127) / 128, 128, 0, s>>>(T, Wt, q_a, level_first, level_count, next_first, q_b, cap, max_leaf, ctr + 1);
    uint32_t h[2];
    B2RT_CUDA_OK(cudaMemcpyAsync(h, ctr, 8, cudaMemcpyDeviceToHost, s));
    B2RT_CUDA_OK(cudaStreamSynchronize(s));
    if (h[1]) { set_error("gpu bvh build: wide node capacity exceeded"); return B2RT_ERR_INVALID; }
    level_first = next_first; level_count = h[0] - next_first;
    std::swap(q_a, q_b);
    if (wide_levels.size() > 4096) { set_error("gpu bvh build: tree too deep"); return B2RT_ERR_INVALID; }
  }
  const uint32_t n_wide = level_first;
  meta->n_wide_nodes = n_wide;
  lap("collapse");
  // 7. subtree sizes, bottom-up
  for (size_t L = wide_levels.size(); L-- > 0;)
    k_sizes<W><<<(wide_levels[L].second + 255) / 256, 256, 0, s>>>(Wt, wide_levels[L].first, wide_levels[L].second);
  uint32_t total_bytes_sat = 0;
  B2RT_CUDA_OK(cudaMemcpyAsync(&total_bytes_sat, Wt.sub_bytes, 4, cudaMemcpyDeviceToHost, s));
  B2RT_CUDA_OK(cudaStreamSynchronize(s));
  const uint64_t total_est = total_bytes_sat == 0xFFFFFFFFu ? (uint64_t)n_wide * NB + (uint64_t)n * PRIM_BYTES : total_bytes_sat;
  lap("sizes");
  // 8. partition, one launch per subtree level
  const uint32_t stride = (treelet_bytes + 127u) & ~127u;
  uint32_t cap_subtrees = (uint32_t)std::min<uint64_t>(8 * (total_est / treelet_bytes) + 1024, 0x3FFFFFFFull);
  uint8_t* slabs = nullptr;
  B2RT_CUDA_OK(buf.alloc(&slabs, (size_t)cap_subtrees * stride));
  TreeletDesc* descs_tmp = nullptr;
  B2RT_CUDA_OK(buf.alloc(&descs_tmp, cap_subtrees));
  uint2* prim_dest = nullptr;
  B2RT_CUDA_OK(buf.alloc(&prim_dest, n));
  uint32_t *r_a = nullptr, *r_b = nullptr;
  B2RT_CUDA_OK(buf.alloc(&r_a, cap_subtrees)); B2RT_CUDA_OK(buf.alloc(&r_b, cap_subtrees));
  { const uint32_t zero = 0; B2RT_CUDA_OK(cudaMemcpyAsync(r_a, &zero, 4, cudaMemcpyHostToDevice, s)); }
  meta->levels.clear();
  uint32_t first_id = 0, n_roots = 1;
  while (n_roots) {
    if (meta->levels.size() >= MAX_LEVELS) { set_error("too many subtree levels; increase treelet_bytes"); return B2RT_ERR_INVALID; }
    meta->levels.push_back(LevelRange{first_id, n_roots});
    B2RT_CUDA_OK(cudaMemsetAsync(ctr + 2, 0, 4, s));
    PartParams P;
    P.slabs = slabs; P.stride = stride; P.descs = descs_tmp; P.roots = r_a; P.n_roots = n_roots; P.first_id = first_id;
    P.next_roots = r_b; P.next_count = ctr + 2; P.next_first_id = first_id + n_roots; P.cap_subtrees = cap_subtrees;
    P.prim_dest = prim_dest; P.budget = treelet_bytes; P.depth_limit = depth_limit; P.node_limit = node_limit; P.overflow = ctr + 1;
    k_partition<W><<<(n_roots + PART_WARPS - 1) / PART_WARPS, PART_WARPS * 32, 0, s>>>(T, Wt, P);
    uint32_t h[2];
    B2RT_CUDA_OK(cudaMemcpyAsync(h, ctr + 1, 8, cudaMemcpyDeviceToHost, s));
    B2RT_CUDA_OK(cudaStreamSynchronize(s));
    if (h[0]) { set_error("gpu bvh build: subtree capacity exceeded"); return B2RT_ERR_INVALID; }
    first_id += n_roots; n_roots = h[1];
    std::swap(r_a, r_b);
  }
  const uint32_t n_subtrees = first_id;
  lap("partition");
  // 9. dense blob
  std::vector<TreeletDesc> descs(n_subtrees);
  B2RT_CUDA_OK(cudaMemcpyAsync(descs.data(), descs_tmp, (size_t)n_subtrees * sizeof(TreeletDesc), cudaMemcpyDeviceToHost, s));
  B2RT_CUDA_OK(cudaStreamSynchronize(s));
  uint64_t total = 0; uint32_t max_bytes = 0;
  for (auto& d : descs) {
    total = (total + 127) & ~127ull;
    if (total / 16 > 0xFFFFFFFFull) { set_error("BVH blob too large"); return B2RT_ERR_INVALID; }
    d.offset16 = (uint32_t)(total / 16);
    total += d.bytes;
    max_bytes = std::max(max_bytes, d.bytes);
  }
  const uint64_t blob_bytes = ((total + 127) & ~127ull) + 128;
  if (dev->blob_cap < blob_bytes) {
    cudaFree(dev->blob); dev->blob = nullptr; dev->blob_cap = 0;
    B2RT_CUDA_OK(cudaMallocAsync(&dev->blob, blob_bytes, s));   // pool memory; free_bvh's cudaFree accepts it
    dev->blob_cap = blob_bytes;
  }
  if (dev->treelet_cap < n_subtrees) {
    cudaFree(dev->treelets); dev->treelets = nullptr; dev->treelet_cap = 0;
    B2RT_CUDA_OK(cudaMallocAsync(&dev->treelets, (size_t)n_subtrees * sizeof(TreeletDesc), s));
    dev->treelet_cap = n_subtrees;
  }
  B2RT_CUDA_OK(cudaMemcpyAsync(dev->treelets, descs.data(), (size_t)n_subtrees * sizeof(TreeletDesc), cudaMemcpyHostToDevice, s));
  k_compact<<<n_subtrees, 128, 0, s>>>(slabs, stride, dev->treelets, NB, dev->blob);
  k_copy_prims<<<(n + 255) / 256, 256, 0, s>>>(geom, sorted, prim_dest, dev->treelets, n, NB, dev->blob);
  B2RT_CUDA_OK(cudaStreamSynchronize(s));
  B2RT_CUDA_OK(cudaGetLastError());
  lap("blob");
  dev->n_treelets = n_subtrees; dev->n_levels = (uint32_t)meta->levels.size(); dev->width = W;
  dev->max_treelet_bytes = max_bytes; dev->blob_bytes = blob_bytes;
  for (uint32_t i = 0; i < dev->n_levels; ++i) dev->levels[i] = meta->levels[i];
  meta->width = W; meta->n_levels = dev->n_levels; meta->max_treelet_bytes = max_bytes;
  meta->treelets = std::move(descs);
  meta->blob.clear();
  meta->build_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
  return B2RT_OK;
}

int build_wide_bvh_device(const b2rt_scene_desc* sc, uint32_t max_leaf, uint32_t width, uint32_t treelet_bytes, cudaStream_t s,
                          DeviceBVH* dev, WideBVH* meta, void* geom_out) {
  if (!sc) { set_error("scene desc is null"); return B2RT_ERR_INVALID; }
  if (sc->n_tris && !sc->tri_verts) { set_error("tri_verts is null"); return B2RT_ERR_INVALID; }
  if (sc->n_spheres && !sc->spheres) { set_error("spheres is null"); return B2RT_ERR_INVALID; }
  const uint64_t n = (uint64_t)sc->n_tris + sc->n_spheres;
  if (width == 0) width = 4;
  if (width != 4 && width != 8) { set_error("the device builder makes 4- or 8-wide trees (widths 2 and 16: host builder)"); return B2RT_ERR_INVALID; }
  // default leaf size of the device builder: 3 (measured against 2 / 4 / 6 / 8 and against the host builder's SAH-cost leaf
  // termination applied to the PLOC tree's own splits, on cfg2, the cfg3 / cfg4 stand-ins and the 10 M soup:
  // profiles/r02_sweep_device_leaf.txt)
  if (max_leaf == 0) max_leaf = 3;
  if (max_leaf > 64) { set_error("max_leaf_size must be <= 64"); return B2RT_ERR_INVALID; }
  if (n == 0) { set_error("gpu bvh build: empty scene"); return B2RT_ERR_INVALID; }
  if (n >= (1u << 29)) { set_error("gpu bvh build: too many primitives"); return B2RT_ERR_INVALID; }
  const uint32_t min_budget = node_bytes(width) + width * max_leaf * PRIM_BYTES;
  if (treelet_bytes == 0) treelet_bytes = (n >= 65536 ? 24 : 20) * 1024;
  treelet_bytes = std::max(treelet_bytes, min_budget) & ~15u;
  if (treelet_bytes > 160 * 1024) { set_error("treelet_bytes exceeds the shared-memory budget (160 KiB)"); return B2RT_ERR_INVALID; }
  return width == 8 ? build_device_w<8>(*sc, max_leaf, treelet_bytes, s, dev, meta, (float4*)geom_out)
                    : build_device_w<4>(*sc, max_leaf, treelet_bytes, s, dev, meta, (float4*)geom_out);
}

}  // namespace b2rt
