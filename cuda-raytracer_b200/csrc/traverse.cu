// Breadth-first wide-BVH traversal with dynamic ray scheduling, sm_100a.
//
// Replaces the reference's per-level machinery:
//   kernelScanCounts                      src/cudaRenderer.cu:1317-1431  -> k_schedule_level
//   kernelRayIntersectSingle/Level        src/cudaRenderer.cu:1304-1310, 1435-1489, 846-1297 -> k_traverse
//   sharedMemExclusiveScan                src/exclusiveScan.cu_inl:73-110 -> ballot/popc + shuffle scans
//   kernelClearIntersections/Merge        src/cudaRenderer.cu:490-540    -> packed (t, prim) 64-bit atomicMin
// Design (DESIGN.md "Traversal"): the BVH is cut into subtrees that fit in shared memory.  One pass
// per subtree LEVEL: rays are grouped by the subtree they must visit; a CTA owns one subtree at a
// time, stages its blob (SoA wide nodes + 48-byte primitive records) with ONE TMA bulk copy
// (cp.async.bulk + mbarrier).  Level 0 (the root subtree, visited by every ray) reads the caller's DENSE SoA ray list
// (origin, direction, hit word): every warp claims runs of consecutive rays from a global cursor and hands them to its
// lanes as they go idle (coalesced streaming loads).  Deeper levels read ray ids grouped by subtree and gather the
// three 8/16-byte records.  A lane walks its ray through the subtree with a per-thread stack in shared memory; leaf
// visits are decoupled: (lane, primitive) items go to a per-warp queue that the whole warp drains 32 at a time.  Rays that
// leave through an EXIT child are pushed as (child subtree, ray id) pairs: warp ballot/popc exclusive
// scan into a per-warp staging ring, one global atomicAdd per flush.  Between levels a single-CTA
// scan turns per-subtree counts into segment offsets + a chunk work list and a scatter kernel
// regroups the ray ids by subtree.  (Scattering whole 44-byte ray records so that deeper levels could stream too
// was measured: the gather moved into the scatter kernel and cost more than it saved.)  No host round trip
// anywhere: all counts live on the device and every grid is persistent.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>

#include "rt_device.cuh"
#include "traverse.cuh"

namespace b2rt {

namespace {

#ifndef B2RT_OCC4
#define B2RT_OCC4 4
#endif
#ifndef B2RT_TRAV_THREADS
#define B2RT_TRAV_THREADS 256
#endif
constexpr int TRAV_THREADS = B2RT_TRAV_THREADS;
constexpr int TRAV_WARPS = TRAV_THREADS / 32;
constexpr int STAGE_PAIRS = 64;   // per-warp staging ring (flush when > STAGE_PAIRS - 32)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done = 0;
  const uint32_t addr = smem_u32(bar);
  while (!done) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
  }
}
// TMA bulk copy global -> shared, completion signalled on the mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
// explicit shared-state-space accesses by 32-bit address: the per-thread stacks and the ray ring live at run-time
// offsets of dynamic shared memory, and through generic pointers every access paid for rebuilding the generic
// address (S2R SR_CgaCtaId + LEA), about 4 % of the kernel's issue slots
__device__ __forceinline__ uint32_t lds_u32(uint32_t a) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ void sts_u32(uint32_t a, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ float4 lds_f4(uint32_t a) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long lds_u64(uint32_t a) {
  unsigned long long v;
  asm volatile("ld.shared.u64 %0, [%1];" : "=l"(v) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ void sts_u64(uint32_t a, unsigned long long v) { asm volatile("st.shared.u64 [%0], %1;" ::"r"(a), "l"(v) : "memory"); }
__device__ __forceinline__ uint32_t atoms_add(uint32_t a, uint32_t v) {
  uint32_t old;
  asm volatile("atom.shared.add.u32 %0, [%1], %2;" : "=r"(old) : "r"(a), "r"(v) : "memory");
  return old;
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- k_schedule_level: exclusive scan of per-subtree ray counts -> segment offsets + chunk list ----
// One CTA of 1024 threads; warp-shuffle scans (no volatile-smem warp-synchronous code).
__device__ __forceinline__ uint32_t warp_incl_scan(uint32_t v) {
  const uint32_t lane = threadIdx.x & 31;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    uint32_t n = __shfl_up_sync(0xffffffffu, v, d);
    if (lane >= (uint32_t)d) v += n;
  }
  return v;
}
// block-wide exclusive scan of two values at once; returns totals through tot_a/tot_b
__device__ __forceinline__ void block_excl_scan2(uint32_t a, uint32_t b, uint32_t* ea, uint32_t* eb, uint32_t* tot_a,
                                                 uint32_t* tot_b, uint32_t* sh /* 2*32+2 */) {
  const uint32_t lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  uint32_t ia = warp_incl_scan(a), ib = warp_incl_scan(b);
  if (lane == 31) { sh[w] = ia; sh[32 + w] = ib; }
  __syncthreads();
  if (w == 0) {
    uint32_t xa = lane < nw ? sh[lane] : 0, xb = lane < nw ? sh[32 + lane] : 0;
    uint32_t sa = warp_incl_scan(xa), sb = warp_incl_scan(xb);
    sh[lane] = sa - xa; sh[32 + lane] = sb - xb;
    if (lane == 31) { sh[64] = sa; sh[65] = sb; }
  }
  __syncthreads();
  *ea = sh[w] + ia - a; *eb = sh[32 + w] + ib - b;
  *tot_a = sh[64]; *tot_b = sh[65];
  __syncthreads();
}

// One CTA per tile of 1024 subtrees; the tiles' running totals are chained through global memory (tile t waits for the
// inclusive prefix of tile t - 1, adds its own totals, publishes): a level of 45 K subtrees is 45 hops of about a
// microsecond instead of 45 iterations of one CTA (528 us on the 10 M-triangle soup, round 1).  Tiles are handed out by
// a ticket, so a CTA only ever waits for one that is already running.  The last tile clears the flags again: when it
// has seen its predecessor's flag, every earlier flag has been read by ITS successor, so the launch leaves the scratch
// area as it found it (no per-launch argument: the launch can be replayed from a CUDA graph).
// scratch: [0] ticket, [1 + 3 t ...] = (flag, ray prefix, chunk prefix) of tile t.
__global__ void __launch_bounds__(1024, 1)
k_schedule_level(uint32_t* __restrict__ cnt, uint32_t* __restrict__ seg_off, uint32_t* __restrict__ cursor,
                 uint4* __restrict__ chunks, uint32_t* __restrict__ ctrl, uint32_t first, uint32_t n, uint32_t chunk_rays,
                 uint32_t chunk_cap, uint32_t level, uint32_t pair_cap, uint32_t want_chunks, uint32_t chunk_min,
                 uint32_t* __restrict__ scratch) {
  // A level with few rays gets smaller chunks, so that the launch still has `want_chunks` of them (a few per resident
  // CTA): with 1024-ray chunks a launch of 1 M rays is 1000 chunks for 592 CTAs and its second round runs on a
  // half-empty GPU.  The level's ray count is the pair count of the level above.
  {
    const uint32_t total = min(ctrl[(level - 1u) & 1u], pair_cap);
    const uint32_t fit = (total / max(want_chunks, 1u) + 127u) & ~127u;
    chunk_rays = min(chunk_rays, max(chunk_min, fit));
  }
  __shared__ uint32_t sh[66];
  __shared__ uint4 big[1024];   // (subtree, seg offset, count, chunk base) of subtrees with many chunks
  __shared__ uint32_t n_big, s_tile, s_run_off, s_run_chunks;
  if (threadIdx.x == 0) { n_big = 0; s_tile = atomicAdd(&scratch[0], 1u); }
  __syncthreads();
  const uint32_t tile = s_tile, n_tiles = gridDim.x;
  const uint32_t i = tile * 1024u + threadIdx.x;
  const uint32_t t = first + i;
  const uint32_t c = i < n ? cnt[t] : 0u;
  if (i < n) cnt[t] = 0;   // self-cleaning: every level >= 1 is scheduled exactly once per trace
  const uint32_t nch = (c + chunk_rays - 1) / chunk_rays;
  uint32_t eo, ec, to, tc;
  block_excl_scan2(c, nch, &eo, &ec, &to, &tc, sh);
  if (threadIdx.x == 0) {
    uint32_t ro = 0, rc = 0;
    if (tile > 0) {
      volatile uint32_t* prev = scratch + 1 + 3 * (tile - 1);
      while (prev[0] == 0u) { }
      __threadfence();
      ro = prev[1]; rc = prev[2];
    }
    volatile uint32_t* mine = scratch + 1 + 3 * tile;
    mine[1] = ro + to; mine[2] = rc + tc;
    __threadfence();
    if (tile + 1 < n_tiles) mine[0] = 1u;
    s_run_off = ro; s_run_chunks = rc;
    if (tile + 1 == n_tiles) {   // the last ticket: every tile has one, so the counter can go back to zero
      scratch[0] = 0;
      for (uint32_t k = 0; k + 1 < n_tiles; ++k) scratch[1 + 3 * k] = 0u;
      uint32_t run_chunks = rc + tc;
      if (run_chunks > chunk_cap) { run_chunks = chunk_cap; ctrl[CTRL_OVERFLOW] = 1; }
      ctrl[CTRL_NCHUNKS] = run_chunks;
      ctrl[CTRL_NEXT0 + (level & 1)] = 0;   // chunk cursor of THIS level's traversal
      ctrl[level & 1] = 0;                  // pair counter the traversal of THIS level appends to
    }
  }
  __syncthreads();
  const uint32_t run_off = s_run_off, run_chunks = s_run_chunks;
  if (i < n) {
    const uint32_t off = run_off + eo, cb = run_chunks + ec;
    seg_off[t] = off; cursor[t] = 0;
    if (nch <= 4) {
      for (uint32_t k = 0; k < nch; ++k)
        if (cb + k < chunk_cap) chunks[cb + k] = make_uint4(t, off + k * chunk_rays, min(chunk_rays, c - k * chunk_rays), 0);
    } else {
      const uint32_t slot = atomicAdd(&n_big, 1u);
      big[slot] = make_uint4(t, off, c, cb);   // at most 1024 subtrees per tile
    }
  }
  __syncthreads();
  const uint32_t nb = min(n_big, 1024u);
  for (uint32_t b = threadIdx.x >> 5; b < nb; b += blockDim.x >> 5) {   // one warp per subtree with many chunks
    const uint4 e = big[b];
    const uint32_t nch_b = (e.z + chunk_rays - 1) / chunk_rays;
    for (uint32_t k = threadIdx.x & 31; k < nch_b; k += 32)
      if (e.w + k < chunk_cap) chunks[e.w + k] = make_uint4(e.x, e.y + k * chunk_rays, min(chunk_rays, e.z - k * chunk_rays), 0);
  }
}

// ---- k_count_*: rays queued per subtree of the next level = histogram of the pair list -------------------
// The traversal used to count with one global atomicAdd per push.  On scenes where every ray is pushed several times
// per level (10 M triangle soup: 2-3 pushes per ray at level 0 onto 570 counters) those atomics saturated the L2
// atomic path and back-pressured the load/store pipe of every SM: the level-0 kernel ran at 9 % issue utilisation
// with 52 % of the warp samples waiting on SHARED-memory loads queued behind them.
// Tiled variant: a CTA histograms all its tiles in shared memory and flushes each non-empty bin once.
__global__ void __launch_bounds__(256)
k_count_tiled(const uint2* __restrict__ pairs, const uint32_t* __restrict__ pair_count, uint32_t* __restrict__ cnt, uint32_t pair_cap,
              uint32_t first, uint32_t K) {
  extern __shared__ uint32_t s_hist[];
  const uint32_t n = min(*pair_count, pair_cap);
  for (uint32_t i = threadIdx.x; i < K; i += 256) s_hist[i] = 0;
  __syncthreads();
  // four pairs per thread and iteration, their loads in flight together
  for (uint32_t i0 = blockIdx.x * 1024 + threadIdx.x; i0 < n; i0 += gridDim.x * 1024) {
    uint32_t t[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) { const uint32_t i = i0 + (uint32_t)k * 256u; t[k] = i < n ? pairs[i].x : 0xFFFFFFFFu; }
#pragma unroll
    for (int k = 0; k < 4; ++k) if (t[k] != 0xFFFFFFFFu) atomicAdd(&s_hist[t[k] - first], 1u);
  }
  __syncthreads();
  for (uint32_t i = threadIdx.x; i < K; i += 256) {
    const uint32_t c = s_hist[i];
    if (c) atomicAdd(&cnt[first + i], c);
  }
}
// Levels with more subtrees than a shared-memory histogram holds: warp-aggregated global atomics (many counters, so
// little contention).
__global__ void __launch_bounds__(256)
k_count(const uint2* __restrict__ pairs, const uint32_t* __restrict__ pair_count, uint32_t* __restrict__ cnt, uint32_t pair_cap) {
  const uint32_t n = min(*pair_count, pair_cap);
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t base = (blockIdx.x * blockDim.x + threadIdx.x) - lane; base < n; base += stride) {
    const uint32_t i = base + lane;
    const uint32_t t = i < n ? pairs[i].x : 0xFFFFFFFFu;
    const bool valid = t != 0xFFFFFFFFu;
    const uint32_t active = __ballot_sync(0xffffffffu, valid);
    if (!valid) continue;
    const uint32_t peers = __match_any_sync(active, t);
    if (lane == (uint32_t)__ffs(peers) - 1u) atomicAdd(&cnt[t], (uint32_t)__popc(peers));
  }
}

// ---- k_scatter: regroup ray ids by subtree (counting-sort scatter, warp-aggregated cursors) ---------
__global__ void __launch_bounds__(256)
k_scatter(const uint2* __restrict__ pairs, const uint32_t* __restrict__ pair_count, const uint32_t* __restrict__ seg_off,
          uint32_t* __restrict__ cursor, uint32_t* __restrict__ ids_sorted, uint32_t pair_cap) {
  const uint32_t n = min(*pair_count, pair_cap);
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t stride = gridDim.x * blockDim.x;
  // whole warps iterate together so the match/shuffle masks are well defined
  for (uint32_t base = (blockIdx.x * blockDim.x + threadIdx.x) - lane; base < n; base += stride) {
    uint32_t i = base + lane;
    bool valid = i < n;
    uint2 p = valid ? pairs[i] : make_uint2(0xFFFFFFFFu, 0u);
    uint32_t active = __ballot_sync(0xffffffffu, valid);
    if (!valid) continue;
    uint32_t peers = __match_any_sync(active, p.x);
    uint32_t leader = __ffs(peers) - 1;
    uint32_t rank = __popc(peers & ((1u << lane) - 1u));
    uint32_t basepos = 0;
    if (lane == leader) basepos = atomicAdd(&cursor[p.x], (uint32_t)__popc(peers));
    basepos = __shfl_sync(peers, basepos, leader);
    ids_sorted[seg_off[p.x] + basepos + rank] = p.y;
  }
}

// Tiled variant: a CTA ranks a tile of pairs in a shared-memory histogram over the level's subtrees and then
// reserves ONE global cursor range per (tile, subtree) instead of one per warp-group -- the global atomics on the
// few hot cursors of a small level (CBbunny: 119 subtrees) were the cost of the plain kernel.
#ifndef B2RT_SCATTER_TILE
#define B2RT_SCATTER_TILE 2048
#endif
constexpr int SCATTER_TILE = B2RT_SCATTER_TILE;   // pairs per CTA iteration (8 per thread)
__global__ void __launch_bounds__(256)
k_scatter_tiled(const uint2* __restrict__ pairs, const uint32_t* __restrict__ pair_count, const uint32_t* __restrict__ seg_off,
                uint32_t* __restrict__ cursor, uint32_t* __restrict__ ids_sorted, uint32_t pair_cap, uint32_t first, uint32_t K) {
  extern __shared__ uint32_t s_hist[];   // [K] counts, then [K] bases
  uint32_t* s_cnt = s_hist;
  uint32_t* s_base = s_hist + K;
  const uint32_t n = min(*pair_count, pair_cap);
  for (uint32_t tile = blockIdx.x * SCATTER_TILE; tile < n; tile += gridDim.x * SCATTER_TILE) {
    for (uint32_t i = threadIdx.x; i < K; i += 256) s_cnt[i] = 0;
    __syncthreads();
    uint2 p[SCATTER_TILE / 256];
    uint32_t r[SCATTER_TILE / 256];
#pragma unroll
    for (int j = 0; j < SCATTER_TILE / 256; ++j) {
      const uint32_t idx = tile + j * 256 + threadIdx.x;
      p[j] = idx < n ? pairs[idx] : make_uint2(0xFFFFFFFFu, 0u);
      if (idx < n) r[j] = atomicAdd(&s_cnt[p[j].x - first], 1u);
    }
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < K; i += 256) {
      const uint32_t c = s_cnt[i];
      if (c) s_base[i] = seg_off[first + i] + atomicAdd(&cursor[first + i], c);
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < SCATTER_TILE / 256; ++j)
      if (p[j].x != 0xFFFFFFFFu) ids_sorted[s_base[p[j].x - first] + r[j]] = p[j].y;
    __syncthreads();
  }
}

// ---- k_traverse -----------------------------------------------------------------------------------
struct TravParams {
  const uint8_t* blob;
  const TreeletDesc* treelets;
  // the dense ray list: (origin | tmin), (direction | tmax), packed hit word, indexed by ray id
  const float4* ray_o;
  const float4* ray_d;
  unsigned long long* hits;
  const uint32_t* ids;        // levels >= 1: ray ids grouped by subtree; nullptr at level 0 (chunk ranges ARE ray ids)
  const uint4* chunks;
  uint32_t* ctrl;
  uint32_t* cnt;              // per-subtree counts for the NEXT level
  uint2* pairs;               // (child subtree, ray id) output
  uint32_t pair_cap;
  uint32_t level;
  uint32_t n_treelets, n_rays_cap;
  uint32_t chunk_rays;
  uint32_t chunk0_max;        // upper bound of the level-0 chunk size
  uint32_t stack_off;         // byte offset of the per-thread traversal stacks inside dynamic shared memory
  uint32_t ring_off;          // byte offset of the ray ring
  const uint32_t* n_active;
  TraceCounters* counters;
};

template <int W>
struct NodeView {
  static constexpr int BYTES = (int)node_bytes(W);
  static constexpr int SLOT_BITS = (int)slot_bits(W);
};
constexpr uint32_t STACK_TN_MASK = 0xFFFFF000u;   // stack entry: [31:12] entry distance bits, [11:0] node << SLOT_BITS | slot


// flush one warp's staged pairs: one global reservation, coalesced 8-byte stores, per-subtree counts
__device__ __forceinline__ void flush_pairs(uint32_t stage_addr, uint32_t& n_staged, const TravParams& P, uint32_t lane) {
  uint32_t n = n_staged;
  uint32_t base = 0;
  if (lane == 0) base = atomicAdd(&P.ctrl[P.level & 1], n);
  base = __shfl_sync(0xffffffffu, base, 0);
  for (uint32_t k = lane; k < n; k += 32) {
    const unsigned long long p = lds_u64(stage_addr + k * 8u);
    if (base + k < P.pair_cap) {
      P.pairs[base + k] = make_uint2((uint32_t)p, (uint32_t)(p >> 32));   // (per-subtree counts are taken by k_count_* afterwards)
    } else {
      P.ctrl[CTRL_OVERFLOW] = 1;
    }
  }
  __syncwarp();
  n_staged = 0;
}

// Closest-hit merge.  At level 0 a ray is visited exactly once and nothing else touches its hit word, so the improved
// word is simply stored; deeper levels can hold the same ray in several subtrees at once and merge with the packed
// (t, prim) 64-bit atomicMin.
__device__ __forceinline__ void retire_hit(const TravParams& P, uint32_t rid, unsigned long long word) {
  if (P.level == 0) P.hits[rid] = word;
  else atomicMin(&P.hits[rid], word);
}

#ifdef B2RT_CHECKS
#define B2_CHECK(cond, code, info) do { if (!(cond)) { atomicExch(&P.ctrl[6], (uint32_t)(code)); atomicExch(&P.ctrl[7], (uint32_t)(info)); } } while (0)
#else
#define B2_CHECK(cond, code, info) do { } while (0)
#endif
constexpr uint32_t REF_NONE = 0xFFFFFFFEu;   // "no current node" marker of the traversal loop (tag EMPTY)
#ifndef B2RT_MIN_IDLE
#define B2RT_MIN_IDLE 20
#endif
constexpr int REFILL_MIN_IDLE = B2RT_MIN_IDLE;   // refill a warp's idle lanes once this many are idle

__device__ __forceinline__ float fast_rcp(float x) {
  float r;
  asm("rcp.approx.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
// three-input fp32 min / max (sm_100: one FMNMX3 instead of two FMNMX)
__device__ __forceinline__ float fmax3(float a, float b, float c) {
  float r;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
  return r;
}
__device__ __forceinline__ float fmin3(float a, float b, float c) {
  float r;
  asm("min.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
  return r;
}

// Leaf work is DECOUPLED from the per-lane tree walk (round 2).  A lane that reaches a leaf does not test its
// primitives itself: it appends (lane, primitive) items to a per-warp queue in shared memory and goes on walking.
// Whenever 32 items are queued the whole warp drains them, one ray-primitive test per lane (the ray is fetched from
// the owning lane with shuffles, the result is merged into the owner's packed (t, prim) word in shared memory with a
// 64-bit atomicMin).  The node phase and the primitive phase therefore each run with (nearly) full warps; before, a
// warp executed the node code and the leaf loop back to back for a handful of lanes each (19.7 of 32 lanes active per
// instruction, profiles/r01_traverse_ncu.md).  The closest hit is the argmin of (t, prim) over every primitive tested,
// so the result does not depend on the order of the tests; between a leaf visit and its drain a lane culls with a
// stale (larger) best_t, which only adds visits.
#ifndef B2RT_LEAF_TAKE
#define B2RT_LEAF_TAKE 3
#endif
// Items wait in the queue until DRAIN_MIN are there (measured: 32 = full batches only is best; the owners cull with a
// stale best_t meanwhile, which costs about 1 % more primitive tests).
#ifndef B2RT_DRAIN_MIN
#define B2RT_DRAIN_MIN 32
#endif
constexpr uint32_t DRAIN_MIN = B2RT_DRAIN_MIN;
constexpr uint32_t LEAF_TAKE = B2RT_LEAF_TAKE;          // primitives a lane queues per iteration (1..4)
constexpr uint32_t QCAP = 32 + 32 * LEAF_TAKE;          // < 32 left over + one round of appends
constexpr uint32_t ITEM_PRIM_MASK = 0x07FFFFFFu;        // item: [31:27] owning lane, [26:0] primitive index in the blob
#ifndef B2RT_GRAB0
#define B2RT_GRAB0 256
#endif
constexpr uint32_t GRAB0 = B2RT_GRAB0;                  // level 0: rays a warp claims from the global cursor at a time
// Per-warp scratch in shared memory.  Everything a warp touches in the traversal loop hangs off ONE shared-space base
// address (wb) plus immediates: through C++ pointers / arrays every access re-derived its address (S2UR SR_CgaCtaId +
// ULEA ... per access, 6 % of the issue slots of the first version of this kernel).
struct WarpLocal {
  unsigned long long stage[STAGE_PAIRS];   // (child subtree | ray id << 32) pairs waiting for a flush
  uint32_t items[QCAP];                    // primitive-test queue
  unsigned long long best[32];             // packed (t, prim) of the ray each lane holds
  float4 ray_o[32];                        // origin | t_min of the ray each lane holds (read by whichever lane tests
  float4 ray_d[32];                        // one of its primitives; the owner itself walks with inv / -o*inv only)
};
constexpr uint32_t WL_STAGE = 0, WL_ITEMS = STAGE_PAIRS * 8, WL_BEST = WL_ITEMS + QCAP * 4, WL_RO = WL_BEST + 32 * 8, WL_RD = WL_RO + 512;
static_assert(sizeof(WarpLocal) == WL_RD + 512 && sizeof(WarpLocal) % 16 == 0 && WL_RO % 16 == 0, "WarpLocal layout");

__device__ __forceinline__ void sts_f4(uint32_t a, float4 v) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
// streaming loads of the level-0 ray list (read exactly once)
__device__ __forceinline__ float4 ldg_stream_f4(const float4* p) {
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}

template <int W> constexpr int STACK_ENTRIES_V = (int)stack_entries(W);   // (a constant the device code may name)
template <int W, bool ANYHIT, bool STATS>
__global__ void __launch_bounds__(TRAV_THREADS, (W <= 4 ? B2RT_OCC4 : (W == 8 ? 2 : 1)))
k_traverse(const TravParams P) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ __align__(8) uint64_t s_bar;
  __shared__ uint4 s_chunk;
  __shared__ uint32_t s_next_ray;
  __shared__ __align__(16) WarpLocal s_w[TRAV_WARPS];

  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t lane_lt = (1u << lane) - 1u;
  constexpr uint32_t NB = NodeView<W>::BYTES;
  constexpr int SLOT_BITS = NodeView<W>::SLOT_BITS;
  // The three base addresses below go through an opaque asm move: left to itself ptxas re-derives each of them at
  // every use (S2R SR_CgaCtaId + MOV + LEA + ..., 8 instructions for the stack address of every pop) instead of
  // keeping them in a register.
  uint32_t sbase, stack, wb;
  asm volatile("mov.u32 %0, %1;" : "=r"(sbase) : "r"(smem_u32(smem)));   // the staged subtree (nodes, then primitives)
  // per-thread stack in shared memory, entry k of thread t at word k * TRAV_THREADS + t (conflict-free)
  asm volatile("mov.u32 %0, %1;" : "=r"(stack) : "r"(sbase + P.stack_off + threadIdx.x * 4u));
  asm volatile("mov.u32 %0, %1;" : "=r"(wb) : "r"(smem_u32(&s_w[warp])));   // this warp's scratch
  uint32_t cur_treelet = 0xFFFFFFFFu;
  uint32_t phase = 0;
  uint32_t n_staged = 0;   // warp-uniform
  uint32_t qn = 0;         // warp-uniform copy of s_w[warp].qn
  unsigned long long st_nodes = 0, st_prims = 0, st_visits = 0, st_push = 0, st_upd = 0;

  if (threadIdx.x == 0) mbar_init(&s_bar, 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncthreads();
  // ---- level 0: every active ray visits the root subtree; the dense ray list is the work list.  A warp claims GRAB0
  // consecutive rays at a time from a global cursor and hands them to its lanes as they go idle (coalesced streaming
  // loads), so the only barrier of the launch is the one at its end.  (Round 1 streamed the list through a CTA-wide
  // shared-memory ring with TMA bulk copies; its bookkeeping -- slot accounting, mbarrier phases, re-arming -- cost
  // more issue slots per refill than the load latency it hid, and its 10 KB now hold the lanes' ray records.)
  const uint32_t n_root = P.level == 0 ? *P.n_active : 0u;
  uint32_t w_next = 0, w_end = 0;   // warp-uniform: the warp's claimed range of the level-0 list
  const uint32_t n_chunks = P.level == 0 ? 1u : P.ctrl[CTRL_NCHUNKS];
  bool stream_started = false;
  for (;;) {
    if (threadIdx.x == 0) {
      uint4 ch = make_uint4(0xFFFFFFFFu, 0, 0, 0);
      if (P.level == 0) {
        if (!stream_started) ch = make_uint4(0u, 0u, 0xFFFFFFFFu, 0u);   // the one pseudo-chunk of level 0: the whole ray list
      } else {
        const uint32_t c = atomicAdd(&P.ctrl[CTRL_NEXT0 + (P.level & 1)], 1u);
        if (c < n_chunks) ch = P.chunks[c];
      }
      s_chunk = ch;
      s_next_ray = 0;
    }
    stream_started = true;
    __syncthreads();
    const uint4 chunk = s_chunk;
    if (chunk.x == 0xFFFFFFFFu) break;
    const TreeletDesc td = P.treelets[chunk.x];
    if (chunk.x != cur_treelet) {
      if (threadIdx.x == 0) {
        fence_proxy_async();
        mbar_expect_tx(&s_bar, td.bytes);
        bulk_g2s(smem, P.blob + (size_t)td.offset16 * 16, td.bytes, &s_bar);
      }
      mbar_wait(&s_bar, phase);
      phase ^= 1;
      cur_treelet = chunk.x;
      if (STATS && threadIdx.x == 0) atomicAdd(&P.counters->staged_bytes, (unsigned long long)td.bytes);
    }
    const uint32_t prims_addr = sbase + td.n_nodes * NB;   // shared-space address of the primitive records

    // ---- per-lane ray state; lanes are refilled from the chunk as their rays finish -------------------
    uint32_t rid = 0;
    float best_t = 0.f;
    uint32_t h0_t = 0, h0_id = 0xFFFFFFFFu;   // the ray's hit word when the lane took it (retire writes only if it improved)
    f3 inv = mk3(0, 0, 0), noi = mk3(0, 0, 0);
    float tmin = 0.f;
    uint32_t nx = 0, ny = 0, nz = 0;   // byte offsets of the near plane rows (far rows: 12W - nx, 20W - ny, 28W - nz)
    int sp = 0;
    uint32_t cur = REF_NONE;
    bool have = false;
    bool exhausted = false;   // warp-uniform: the chunk has no more rays to hand out
    uint32_t idle_min = (uint32_t)REFILL_MIN_IDLE;   // idle lanes that trigger a refill; 32 (= leave the loop) once exhausted

    // one batch of the warp's queue: items [head, head + n), one ray-primitive test per lane
    auto drain = [&](uint32_t head, uint32_t n) {
      if (lane < n) {
        const uint32_t item = lds_u32(wb + WL_ITEMS + (head + lane) * 4u);
        const uint32_t src = item >> 27;
        const uint32_t pa = prims_addr + (item & ITEM_PRIM_MASK) * (uint32_t)PRIM_BYTES;
        B2_CHECK((item & ITEM_PRIM_MASK) < td.n_prims, 4, item);
        const float4 ro = lds_f4(wb + WL_RO + src * 16u), rd = lds_f4(wb + WL_RD + src * 16u);
        PrimRec p;
        p.a = lds_f4(pa); p.b = lds_f4(pa + 16u); p.c = lds_f4(pa + 32u);
        if (STATS) st_prims++;
        float t, u, v;
        const uint32_t pid = __float_as_uint(p.c.y);
        const f3 io = mk3(ro.x, ro.y, ro.z), id = mk3(rd.x, rd.y, rd.z);
        // the upper end of the ray's interval is the owner's packed word: it starts at (tmax, none) and only decreases
        const bool h = (__float_as_uint(p.c.z) != 0u) ? hit_sphere(p, io, id, ro.w, __builtin_huge_valf(), &t)
                                                      : hit_triangle(p, io, id, ro.w, __builtin_huge_valf(), &t, &u, &v);
        if (h) {
          const unsigned long long cand = pack_hit(t, pid);
          if (cand < lds_u64(wb + WL_BEST + src * 8u)) atomicMin(&s_w[warp].best[src], ANYHIT ? pack_hit(0.0f, pid) : cand);
        }
      }
      __syncwarp();
    };
    // after a drain: every lane re-reads its ray's packed word
    auto refresh_best = [&]() {
      if (ANYHIT) {
        const unsigned long long nb = lds_u64(wb + WL_BEST + lane * 8u);
        best_t = __uint_as_float((uint32_t)(nb >> 32));
        if ((uint32_t)nb != 0xFFFFFFFFu) { sp = 0; cur = REF_NONE; }   // occluded: the ray is done
      } else {
        best_t = __uint_as_float(lds_u32(wb + WL_BEST + lane * 8u + 4u));
      }
    };
    auto drain_all = [&]() {
      if (qn == 0) return;
      while (qn) {
        const uint32_t n = min(qn, 32u);
        qn -= n;
        drain(qn, n);
      }
      __syncwarp();
    };

    for (;;) {
      // ---- pop: lanes without a current reference take the nearest entry of their stack that can still matter
      if (cur == REF_NONE) {
        while (sp > 0) {
          const uint32_t e = lds_u32(stack + (uint32_t)(--sp) * (TRAV_THREADS * 4u));
          if (__uint_as_float(e & STACK_TN_MASK) <= best_t) {
            cur = lds_u32(sbase + ((e & 0xFFFu) >> SLOT_BITS) * NB + 24u * W + (e & (uint32_t)(W - 1)) * 4u);
            break;
          }
        }
      }
      const uint32_t m_idle = __ballot_sync(0xffffffffu, cur == REF_NONE);
      if ((uint32_t)__popc(m_idle) >= idle_min) {   // idle_min = 32 once the chunk is handed out
        if (exhausted) break;
        // finish the queued tests (the idle lanes' rays may still have items pending), retire the finished rays,
        // then hand new rays to the idle lanes
        drain_all();
        const bool idle = cur == REF_NONE;
        if (idle && have) {
          const unsigned long long nb = lds_u64(wb + WL_BEST + lane * 8u);
          if ((uint32_t)(nb >> 32) != h0_t || (uint32_t)nb != h0_id) { retire_hit(P, rid, nb); if (STATS) st_upd++; }
          have = false;
        }
        if (!idle) refresh_best();
        const uint32_t n_idle = __popc(m_idle), rank = __popc(m_idle & lane_lt);
        bool take;
        float4 ro = make_float4(0.f, 0.f, 0.f, 0.f), rd = make_float4(0.f, 0.f, 1.f, 0.f);
        unsigned long long h = 0;
        if (P.ids) {
          // levels >= 1: ray ids grouped by subtree; the CTA's warps share the chunk through a shared-memory cursor
          uint32_t base = 0;
          if (lane == 0) base = atomicAdd(&s_next_ray, n_idle);
          base = __shfl_sync(0xffffffffu, base, 0);
          const uint32_t chunk_first = s_chunk.y, chunk_count = s_chunk.z;
          if (base + n_idle >= chunk_count) { exhausted = true; idle_min = 32u; }
          take = idle && base + rank < chunk_count;
          if (take) {
            rid = P.ids[chunk_first + base + rank];
            B2_CHECK(rid < P.n_rays_cap, 1, rid);
            if (rid >= P.n_rays_cap) rid = 0;
            ro = P.ray_o[rid]; rd = P.ray_d[rid]; h = P.hits[rid];
          }
        } else {
          // level 0: the warp's claimed range of the dense list, re-claimed from the global cursor when it runs out
          if (w_next >= w_end) {
            uint32_t g = 0;
            if (lane == 0) g = atomicAdd(&P.ctrl[CTRL_NEXT0], GRAB0);
            g = __shfl_sync(0xffffffffu, g, 0);
            if (g < n_root) { w_next = g; w_end = (n_root - g < GRAB0) ? n_root : g + GRAB0; }
            else { exhausted = true; idle_min = 32u; }
          }
          const uint32_t avail = w_end - w_next;
          take = idle && rank < avail;
          if (take) {
            rid = w_next + rank;
            ro = ldg_stream_f4(P.ray_o + rid); rd = ldg_stream_f4(P.ray_d + rid); h = __ldcs(P.hits + rid);
          }
          w_next += min(n_idle, avail);
        }
        if (take) {
          sts_f4(wb + WL_RO + lane * 16u, ro); sts_f4(wb + WL_RD + lane * 16u, rd);
          tmin = ro.w;
          h0_t = (uint32_t)(h >> 32); h0_id = (uint32_t)h;
          best_t = __uint_as_float(h0_t);
          sts_u64(wb + WL_BEST + lane * 8u, h);
          // reciprocal direction for the slab test; |d_k| < 1e-18 (incl. +-0) is clamped so that o_k * inv_k
          // stays finite: the ray is then parallel to the slab and the test reduces to lo_k <= o_k <= hi_k
          // (MUFU.RCP, 1 ulp: the slab test only has to be conservative, and the box padding + the 4-ulp slack on
          //  t_far cover it; the IEEE-rounded reciprocal cost 3 x 8 instructions per ray)
          inv = mk3(fabsf(rd.x) > 1e-18f ? fast_rcp(rd.x) : copysignf(1e18f, rd.x),
                    fabsf(rd.y) > 1e-18f ? fast_rcp(rd.y) : copysignf(1e18f, rd.y),
                    fabsf(rd.z) > 1e-18f ? fast_rcp(rd.z) : copysignf(1e18f, rd.z));
          noi = mk3(-(ro.x * inv.x), -(ro.y * inv.y), -(ro.z * inv.z));
          nx = inv.x >= 0.f ? 0u : 12u * W;
          ny = inv.y >= 0.f ? 4u * W : 16u * W;
          nz = inv.z >= 0.f ? 8u * W : 20u * W;
          have = true; sp = 0;
          // not traced: an any-hit ray that is already occluded; a ray whose interval is empty (the renderer's null entries)
          const bool skip = ANYHIT ? h0_id != 0xFFFFFFFFu : best_t < tmin;
          cur = skip ? REF_NONE : 0u;   // INTERNAL node 0 = subtree root
          if (STATS) st_visits++;
        }
        __syncwarp();
      }

      // ---- node phase: every lane whose reference is a wide node tests its W child boxes ------------------
      if ((cur >> 30) == REF_INTERNAL) {
        if (STATS) st_nodes++;
        B2_CHECK((cur & 0x3FFFFFFFu) < td.n_nodes, 2, cur);
        const uint32_t na = sbase + (cur & 0x3FFFFFFFu) * NB;
        const uint32_t fx = 12u * W - nx, fy = 20u * W - ny, fz = 28u * W - nz;
        uint32_t keys[W];
        // sign-ordered slab test: per axis the near plane row is lo (inv >= 0) or hi (inv < 0), chosen once
        // per ray (row offsets nx/ny/nz, fx/fy/fz), so a box costs 6 fma + 3 max + 3 min.  Empty slots hold
        // inverted infinite boxes (lo = +inf, hi = -inf) => t_near = +inf, t_far = -inf => never hit.
        if (W >= 4) {
#pragma unroll
          for (int q = 0; q < W / 4; ++q) {
            const float4 ax = lds_f4(na + nx + 16 * q), bx = lds_f4(na + fx + 16 * q);
            const float4 ay = lds_f4(na + ny + 16 * q), by = lds_f4(na + fy + 16 * q);
            const float4 az = lds_f4(na + nz + 16 * q), bz = lds_f4(na + fz + 16 * q);
            const float axa[4] = {ax.x, ax.y, ax.z, ax.w}, aya[4] = {ay.x, ay.y, ay.z, ay.w}, aza[4] = {az.x, az.y, az.z, az.w};
            const float bxa[4] = {bx.x, bx.y, bx.z, bx.w}, bya[4] = {by.x, by.y, by.z, by.w}, bza[4] = {bz.x, bz.y, bz.z, bz.w};
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              const float tn = fmaxf(fmax3(__fmaf_rn(axa[c], inv.x, noi.x), __fmaf_rn(aya[c], inv.y, noi.y),
                                           __fmaf_rn(aza[c], inv.z, noi.z)), tmin);
              const float tf = fminf(fmin3(__fmaf_rn(bxa[c], inv.x, noi.x), __fmaf_rn(bya[c], inv.y, noi.y),
                                           __fmaf_rn(bza[c], inv.z, noi.z)), best_t);
              const bool hit = tn <= tf * 1.0000004f;
              // key: entry distance (low byte dropped = rounded down, keeps order for t >= 0) | child slot in the low byte (one PRMT)
              keys[(q * 4 + c) % W] = hit ? __byte_perm(__float_as_uint(tn), (uint32_t)(q * 4 + c), 0x3214) : 0xFFFFFFFFu;
            }
          }
        } else {   // W == 2: 8-byte rows
          const unsigned long long ax = lds_u64(na + nx), bx = lds_u64(na + fx), ay = lds_u64(na + ny), by = lds_u64(na + fy);
          const unsigned long long az = lds_u64(na + nz), bz = lds_u64(na + fz);
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            const int sh = 32 * c;
            const float tn = fmaxf(fmax3(__fmaf_rn(__uint_as_float((uint32_t)(ax >> sh)), inv.x, noi.x), __fmaf_rn(__uint_as_float((uint32_t)(ay >> sh)), inv.y, noi.y),
                                         __fmaf_rn(__uint_as_float((uint32_t)(az >> sh)), inv.z, noi.z)), tmin);
            const float tf = fminf(fmin3(__fmaf_rn(__uint_as_float((uint32_t)(bx >> sh)), inv.x, noi.x), __fmaf_rn(__uint_as_float((uint32_t)(by >> sh)), inv.y, noi.y),
                                         __fmaf_rn(__uint_as_float((uint32_t)(bz >> sh)), inv.z, noi.z)), best_t);
            const bool hit = tn <= tf * 1.0000004f;
            keys[c % W] = hit ? __byte_perm(__float_as_uint(tn), (uint32_t)c, 0x3214) : 0xFFFFFFFFu;
          }
        }
#define B2_CE(a, b) { const uint32_t lo_ = min(keys[(a) % W], keys[(b) % W]), hi_ = max(keys[(a) % W], keys[(b) % W]); keys[(a) % W] = lo_; keys[(b) % W] = hi_; }
        if (W == 2) {
          B2_CE(0, 1)
        } else if (W == 4) {
          B2_CE(0, 1) B2_CE(2, 3) B2_CE(0, 2) B2_CE(1, 3) B2_CE(1, 2)
        } else if (W == 8) {
          B2_CE(0, 1) B2_CE(2, 3) B2_CE(4, 5) B2_CE(6, 7)
          B2_CE(0, 2) B2_CE(1, 3) B2_CE(4, 6) B2_CE(5, 7)
          B2_CE(1, 2) B2_CE(5, 6) B2_CE(0, 4) B2_CE(3, 7)
          B2_CE(1, 5) B2_CE(2, 6)
          B2_CE(1, 4) B2_CE(3, 6)
          B2_CE(2, 4) B2_CE(3, 5)
          B2_CE(3, 4)
        } else {   // W == 16: bitonic network (80 compare-exchanges), fully unrolled
#pragma unroll
          for (int k = 2; k <= W; k <<= 1) {
#pragma unroll
            for (int j = k >> 1; j > 0; j >>= 1) {
#pragma unroll
              for (int i = 0; i < W; ++i) {
                const int l = i ^ j;
                if (l > i) {
                  const uint32_t a_ = keys[i], b_ = keys[l];
                  const bool up = (i & k) == 0;
                  keys[i] = up ? min(a_, b_) : max(a_, b_);
                  keys[l] = up ? max(a_, b_) : min(a_, b_);
                }
              }
            }
          }
        }
#undef B2_CE
        // far-to-near onto the stack; the nearest child becomes the current reference without a stack round trip.
        // An entry names the child by (node, slot); its reference is read from the node when it is popped.
        const uint32_t node_tag = (cur & 0x3FFFFFFFu) << SLOT_BITS;
#pragma unroll
        for (int q = W - 1; q >= 1; --q) {
          if (keys[q] != 0xFFFFFFFFu) {
            B2_CHECK(sp < STACK_ENTRIES_V<W>, 3, sp);
            sts_u32(stack + (uint32_t)sp * (TRAV_THREADS * 4u), (keys[q] & (STACK_TN_MASK | (uint32_t)(W - 1))) | node_tag);
            ++sp;
          }
        }
        cur = keys[0] != 0xFFFFFFFFu ? lds_u32(na + 24u * W + (keys[0] & (uint32_t)(W - 1)) * 4u) : REF_NONE;
      }

      // ---- leaf references: queue (lane, primitive) items ----------------------------------------------------------
      {
        const bool is_leaf = (cur >> 30) == REF_LEAF;
        // ballot / popc prefix sums of the per-lane counts (1..4: three ballots, one per bit)
        const uint32_t m_leaf = __ballot_sync(0xffffffffu, is_leaf);
        if (m_leaf) {
          const uint32_t first = cur & 0x00FFFFFFu, count = ((cur >> 24) & 63u) + 1u;
          const uint32_t take = is_leaf ? min(count, LEAF_TAKE) : 0u;
          const uint32_t b0 = __ballot_sync(0xffffffffu, take & 1u), b1 = __ballot_sync(0xffffffffu, take & 2u);
          uint32_t pre = __popc(b0 & lane_lt) + 2u * __popc(b1 & lane_lt), tot = __popc(b0) + 2u * __popc(b1);
          if (LEAF_TAKE >= 4) {
            const uint32_t b2 = __ballot_sync(0xffffffffu, take & 4u);
            pre += 4u * __popc(b2 & lane_lt); tot += 4u * __popc(b2);
          }
          B2_CHECK(qn + tot <= QCAP, 6, qn);
          if (is_leaf) {
            const uint32_t at = wb + WL_ITEMS + (qn + pre) * 4u, item = (lane << 27) | first;
#pragma unroll
            for (uint32_t j = 0; j < LEAF_TAKE; ++j)
              if (j < take) sts_u32(at + j * 4u, item + j);
            cur = count > take ? ((REF_LEAF << 30) | ((count - take - 1u) << 24) | (first + take)) : REF_NONE;
          }
          qn += tot;
          __syncwarp();
        }
      }
      // ---- exits: scheduler push; ballot + popc = exclusive scan of the 0/1 flags inside the warp -------------
      {
        const bool do_push = (cur >> 30) == REF_EXIT;
        const uint32_t m = __ballot_sync(0xffffffffu, do_push);
        if (m) {
          B2_CHECK(n_staged + 32 <= STAGE_PAIRS, 6, n_staged);
          if (do_push) {
            B2_CHECK((cur & 0x3FFFFFFFu) < P.n_treelets, 5, cur);
            sts_u64(wb + WL_STAGE + (n_staged + __popc(m & lane_lt)) * 8u, (unsigned long long)(cur & 0x3FFFFFFFu) | ((unsigned long long)rid << 32));
            cur = REF_NONE;
            if (STATS) st_push++;
          }
          n_staged += __popc(m);
          __syncwarp();
          if (n_staged > STAGE_PAIRS - 32) flush_pairs(wb + WL_STAGE, n_staged, P, lane);
        }
      }
      // ---- primitive phase: drain full batches from the tail of the queue, one test per lane --------------------
      if (qn >= DRAIN_MIN) {
        if (DRAIN_MIN >= 32u) {
          do { qn -= 32u; drain(qn, 32u); } while (qn >= 32u);
        } else {
          do { const uint32_t n = min(qn, 32u); qn -= n; drain(qn, n); } while (qn >= DRAIN_MIN);
        }
        refresh_best();
        __syncwarp();
      }
    }
    // the warp is done with the chunk: finish the queued tests and retire the rays still held by the lanes
    drain_all();
    if (have) {
      const unsigned long long nb = lds_u64(wb + WL_BEST + lane * 8u);
      if ((uint32_t)(nb >> 32) != h0_t || (uint32_t)nb != h0_id) { retire_hit(P, rid, nb); if (STATS) st_upd++; }
    }
    __syncthreads();   // every warp is done with this chunk (s_chunk / s_next_ray / subtree smem reusable)
  }
  if (n_staged) flush_pairs(wb + WL_STAGE, n_staged, P, lane);
  if (STATS) {
#pragma unroll
    for (int dlt = 16; dlt > 0; dlt >>= 1) {
      st_nodes += __shfl_xor_sync(0xffffffffu, st_nodes, dlt);
      st_prims += __shfl_xor_sync(0xffffffffu, st_prims, dlt);
      st_visits += __shfl_xor_sync(0xffffffffu, st_visits, dlt);
      st_push += __shfl_xor_sync(0xffffffffu, st_push, dlt);
      st_upd += __shfl_xor_sync(0xffffffffu, st_upd, dlt);
    }
    if (lane == 0) {
      atomicAdd(&P.counters->node_visits, st_nodes);
      atomicAdd(&P.counters->prim_tests, st_prims);
      atomicAdd(&P.counters->subtree_visits, st_visits);
      atomicAdd(&P.counters->pushes, st_push);
      atomicAdd(&P.counters->hit_updates, st_upd);
    }
  }
}

// ---- distance-sliced tracing: list set-up and advance (Tracer::trace_sliced) ------------------------------------
// CTA-aggregated append (256 threads): one global atomic per CTA; returns the list index of the flagged threads.
// Called by every thread of the CTA.
__device__ __forceinline__ uint32_t cta_append(bool flag, uint32_t* counter, uint32_t* s_warp) {
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t m = __ballot_sync(0xffffffffu, flag);
  if (lane == 0) s_warp[warp] = __popc(m);
  __syncthreads();
  if (warp == 0) {
    const uint32_t c = lane < 8 ? s_warp[lane] : 0u;
    uint32_t incl = c;
#pragma unroll
    for (int dlt = 1; dlt < 8; dlt <<= 1) {
      const uint32_t v = __shfl_up_sync(0xffffffffu, incl, dlt);
      if (lane >= (uint32_t)dlt) incl += v;
    }
    const uint32_t total = __shfl_sync(0xffffffffu, incl, 7);
    uint32_t base = 0;
    if (lane == 0 && total) base = atomicAdd(counter, total);
    base = __shfl_sync(0xffffffffu, base, 0);
    if (lane < 8) s_warp[lane] = base + incl - c;
  }
  __syncthreads();
  const uint32_t r = s_warp[warp] + __popc(m & ((1u << lane) - 1u));
  __syncthreads();
  return r;
}

// entry / exit distance of the ray against the scene box (conservativeness comes from the caller's margin)
__device__ __forceinline__ void clip_scene_box(float4 o, float4 d, const float* bb, float* t_enter, float* t_exit) {
  const float ix = fabsf(d.x) > 1e-18f ? __frcp_rn(d.x) : copysignf(1e18f, d.x);
  const float iy = fabsf(d.y) > 1e-18f ? __frcp_rn(d.y) : copysignf(1e18f, d.y);
  const float iz = fabsf(d.z) > 1e-18f ? __frcp_rn(d.z) : copysignf(1e18f, d.z);
  const float ax = (bb[0] - o.x) * ix, bx = (bb[3] - o.x) * ix;
  const float ay = (bb[1] - o.y) * iy, by = (bb[4] - o.y) * iy;
  const float az = (bb[2] - o.z) * iz, bz = (bb[5] - o.z) * iz;
  *t_enter = fmaxf(fmaxf(fminf(ax, bx), fminf(ay, by)), fminf(az, bz));
  *t_exit = fminf(fminf(fmaxf(ax, bx), fmaxf(ay, by)), fmaxf(az, bz));
}

struct SliceList {
  float4* o; float4* d; unsigned long long* h; uint32_t* map; float* exit; uint32_t* n;
};
struct SliceBox { float v[6]; float margin; };

// First list: every ray that can hit the scene at all, with the interval [tmin, max(tmin, t_enter) + first].
__global__ void __launch_bounds__(256)
k_slice_init(const uint32_t* n_dev, const float4* __restrict__ ray_o, const float4* __restrict__ ray_d,
             const unsigned long long* __restrict__ hits, SliceBox box, float first, bool any_hit, SliceList out) {
  __shared__ uint32_t s_warp[8];
  const uint32_t n = *n_dev;
  const uint32_t per = (n + gridDim.x - 1) / gridDim.x;   // a contiguous range per CTA keeps neighbouring rays neighbours
  const uint32_t lo = blockIdx.x * per, hi = min(n, lo + per);
  for (uint32_t at = lo; at < hi; at += 256) {
    const uint32_t i = at + threadIdx.x;
    bool keep = false;
    float4 o = make_float4(0, 0, 0, 0), d = make_float4(0, 0, 1, 0);
    float t_end = 0.f, t_exit = 0.f;
    if (i < hi) {
      o = ray_o[i]; d = ray_d[i];
      float t_enter;
      clip_scene_box(o, d, box.v, &t_enter, &t_exit);
      t_exit = t_exit + box.margin + 1e-4f * fabsf(t_exit);
      t_enter = t_enter - box.margin - 1e-4f * fabsf(t_enter);
      // dropped only when the comparison is TRUE (a NaN keeps the ray): the ray misses the padded box, the box is
      // behind tmin, or beyond tmax
      const bool miss = (t_enter > t_exit) || (t_exit < o.w) || (t_enter > d.w) || (d.w < o.w);
      const bool done = any_hit && (uint32_t)hits[i] != 0xFFFFFFFFu;
      keep = !miss && !done;
      t_end = fminf(d.w, fmaxf(o.w, t_enter) + first);
    }
    const uint32_t k = cta_append(keep, out.n, s_warp);
    if (keep) {
      out.o[k] = o;
      out.d[k] = make_float4(d.x, d.y, d.z, t_end);
      out.h[k] = pack_hit(t_end, 0xFFFFFFFFu);
      out.map[k] = i;
      out.exit[k] = t_exit;
    }
  }
}

// After a pass: rays with a hit are final (their word goes to the caller's array); the others move on to the next
// slice [end, end + len] unless they have left the scene or reached their own tmax.
__global__ void __launch_bounds__(256)
k_slice_advance(SliceList in, const float4* __restrict__ ray_d, unsigned long long* hits, float len, bool last, SliceList out) {
  __shared__ uint32_t s_warp[8];
  const uint32_t n = *in.n;
  const uint32_t per = (n + gridDim.x - 1) / gridDim.x;
  const uint32_t lo = blockIdx.x * per, hi = min(n, lo + per);
  for (uint32_t at = lo; at < hi; at += 256) {
    const uint32_t j = at + threadIdx.x;
    bool keep = false;
    float4 o = make_float4(0, 0, 0, 0), d = make_float4(0, 0, 1, 0);
    float t_end = 0.f, t_exit = 0.f;
    uint32_t i = 0;
    if (j < hi) {
      const unsigned long long h = in.h[j];
      i = in.map[j];
      if ((uint32_t)h != 0xFFFFFFFFu) {
        hits[i] = h;
      } else if (!last) {
        d = in.d[j];
        t_exit = in.exit[j];
        const float t_user = ray_d[i].w;
        const float t_lo = d.w;                    // the slice just traced ended here
        keep = !(t_lo >= t_user) && !(t_lo > t_exit);
        if (keep) { o = in.o[j]; o.w = t_lo; t_end = fminf(t_user, t_lo + len); }
      }
    }
    if (last) continue;
    const uint32_t k = cta_append(keep, out.n, s_warp);
    if (keep) {
      out.o[k] = o;
      out.d[k] = make_float4(d.x, d.y, d.z, t_end);
      out.h[k] = pack_hit(t_end, 0xFFFFFFFFu);
      out.map[k] = i;
      out.exit[k] = t_exit;
    }
  }
}

}  // namespace


// ---- host side ---------------------------------------------------------------------------------------
int upload_bvh(const WideBVH& h, DeviceBVH* d) {
  // grow-only: a re-upload of a scene of similar size costs two copies, no cudaMalloc / cudaFree
  uint8_t* blob = d->blob; TreeletDesc* tl = d->treelets;
  uint64_t blob_cap = d->blob_cap, tl_cap = d->treelet_cap;
  *d = DeviceBVH();
  d->blob = blob; d->treelets = tl; d->blob_cap = blob_cap; d->treelet_cap = tl_cap;
  d->n_treelets = (uint32_t)h.treelets.size();
  d->n_levels = h.n_levels;
  d->width = h.width;
  d->max_treelet_bytes = h.max_treelet_bytes;
  d->blob_bytes = h.blob.size();
  for (uint32_t i = 0; i < h.n_levels; ++i) d->levels[i] = h.levels[i];
  if (h.treelets.empty()) return B2RT_OK;
  if (d->blob_cap < h.blob.size()) {
    cudaFree(d->blob); d->blob = nullptr; d->blob_cap = 0;
    const size_t cap = h.blob.size() + h.blob.size() / 4;
    B2RT_CUDA_OK(cudaMalloc(&d->blob, cap));
    d->blob_cap = cap;
  }
  if (d->treelet_cap < h.treelets.size()) {
    cudaFree(d->treelets); d->treelets = nullptr; d->treelet_cap = 0;
    const size_t cap = h.treelets.size() + h.treelets.size() / 4 + 16;
    B2RT_CUDA_OK(cudaMalloc(&d->treelets, cap * sizeof(TreeletDesc)));
    d->treelet_cap = cap;
  }
  B2RT_CUDA_OK(cudaMemcpy(d->blob, h.blob.data(), h.blob.size(), cudaMemcpyHostToDevice));
  B2RT_CUDA_OK(cudaMemcpy(d->treelets, h.treelets.data(), h.treelets.size() * sizeof(TreeletDesc), cudaMemcpyHostToDevice));
  // cudaMemcpy from pageable memory may return before the DMA has landed, and the non-blocking work streams do
  // not order against the legacy stream: make the upload visible to every stream before anything traverses it
  B2RT_CUDA_OK(cudaDeviceSynchronize());
  return B2RT_OK;
}

void free_bvh(DeviceBVH* d) {
  if (d->blob) cudaFree(d->blob);
  if (d->treelets) cudaFree(d->treelets);
  *d = DeviceBVH();
}

static size_t stack_bytes(uint32_t width) { return (size_t)stack_entries(width) * TRAV_THREADS * 4; }

template <int W, bool A, bool S>
static int prep_kernel(size_t smem, int* occ) {
  B2RT_CUDA_OK(cudaFuncSetAttribute(k_traverse<W, A, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  B2RT_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(occ, k_traverse<W, A, S>, TRAV_THREADS, smem));
  return B2RT_OK;
}

int Tracer::init(const DeviceBVH& b, uint64_t max_rays_, uint32_t pair_factor) {
  // ray-count dependent buffers (large) are kept when only the BVH changes
  if (pair_factor == 0) pair_factor = 4;
  uint64_t want_pairs = max_rays_ * pair_factor + 65536;
  if (want_pairs > 0xFFFF0000ull) want_pairs = 0xFFFF0000ull;
  // test hook: a small pair list forces the overflow / split-and-retry paths (tests/test_gpu.py)
  if (const char* e = getenv("B2RT_DEBUG_PAIR_CAP")) { long long v = atoll(e); if (v >= 1024) want_pairs = (uint64_t)v; }
  int dev = 0;
  B2RT_CUDA_OK(cudaGetDevice(&dev));
  B2RT_CUDA_OK(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
  if (!pairs || want_pairs != pair_cap) {
    cudaFree(pairs); cudaFree(ids_sorted); pairs = nullptr; ids_sorted = nullptr;
    pair_cap = want_pairs;
    B2RT_CUDA_OK(cudaMalloc(&pairs, pair_cap * sizeof(uint2)));
    B2RT_CUDA_OK(cudaMalloc(&ids_sorted, pair_cap * 4));
  }
  max_rays = max_rays_;
  if (!ctrl) {
    B2RT_CUDA_OK(cudaMalloc(&ctrl, 16 * 4));
    B2RT_CUDA_OK(cudaMalloc(&counters, 2 * sizeof(TraceCounters)));
    B2RT_CUDA_OK(cudaMemset(ctrl, 0, 16 * 4));
    B2RT_CUDA_OK(cudaMemset(counters, 0, 2 * sizeof(TraceCounters)));
  }
  // BVH dependent part (small, grow-only)
  bvh = b;
  // dynamic shared memory: the staged subtree blob, then the per-thread traversal stacks
  stack_off = (std::max<size_t>(bvh.max_treelet_bytes, 1024) + 127) & ~(size_t)127;
  ring_off = 0;
  smem_bytes = stack_off + stack_bytes(bvh.width);
  if (const char* e = getenv("B2RT_CHUNK_RAYS")) { int v = atoi(e); if (v >= 32 && v <= (1 << 20)) chunk_rays = (uint32_t)v; }
  if (const char* e = getenv("B2RT_COUNT_CTAS")) { int v = atoi(e); if (v >= 1 && v <= 32) count_ctas = (uint32_t)v; }
  if (const char* e = getenv("B2RT_SCATTER_CTAS")) { int v = atoi(e); if (v >= 1 && v <= 32) scatter_ctas = (uint32_t)v; }
  if (const char* e = getenv("B2RT_CHUNK_MIN")) { int v = atoi(e); if (v >= 32 && v <= (1 << 20)) chunk_min = (uint32_t)v; }
  if (const char* e = getenv("B2RT_CHUNKS_PER_CTA")) { int v = atoi(e); if (v >= 0 && v <= 64) chunks_per_cta = (uint32_t)v; }
  if (const char* e = getenv("B2RT_CHUNK0_MAX")) { int v = atoi(e); if (v >= 32 && v <= (1 << 24)) chunk0_max = (uint32_t)v; }
  chunk_cap = pair_cap / std::min(chunk_rays, chunk_min) + (uint64_t)bvh.n_treelets + 1024;
  int occ = 1, o2 = 1, o3 = 1, o4 = 1;
  int rc;
#define B2_PREP(WW)                                                             \
  do {                                                                          \
    if ((rc = prep_kernel<WW, false, false>(smem_bytes, &occ))) return rc;      \
    if ((rc = prep_kernel<WW, true, false>(smem_bytes, &o2))) return rc;        \
    if ((rc = prep_kernel<WW, false, true>(smem_bytes, &o3))) return rc;        \
    if ((rc = prep_kernel<WW, true, true>(smem_bytes, &o4))) return rc;         \
  } while (0)
  if (bvh.width == 2) B2_PREP(2);
  else if (bvh.width == 8) B2_PREP(8);
  else if (bvh.width == 16) B2_PREP(16);
  else B2_PREP(4);
#undef B2_PREP
  ctas_per_sm = std::max(1, std::min(std::min(occ, o2), std::min(o3, o4)));
  if (getenv("B2RT_VERBOSE"))
    fprintf(stderr, "b2rt: k_traverse W=%u: %d threads, %zu B dynamic smem (subtree %u + stacks %zu + ring %u), %d CTAs/SM x %d SMs\n",
            bvh.width, TRAV_THREADS, smem_bytes, bvh.max_treelet_bytes, stack_bytes(bvh.width), 0u,
            ctas_per_sm, num_sms);
  B2RT_CUDA_OK(cudaFuncSetAttribute(k_scatter_tiled, cudaFuncAttributeMaxDynamicSharedMemorySize, 12288 * 8));
  B2RT_CUDA_OK(cudaFuncSetAttribute(k_count_tiled, cudaFuncAttributeMaxDynamicSharedMemorySize, 12288 * 4));
  const size_t nt = std::max<uint32_t>(1, bvh.n_treelets);
  if (nt_cap < nt) {
    cudaFree(cnt); cudaFree(seg_off); cudaFree(cursor); cnt = seg_off = cursor = nullptr;
    nt_cap = nt + nt / 4 + 16;
    B2RT_CUDA_OK(cudaMalloc(&cnt, nt_cap * 4));
    B2RT_CUDA_OK(cudaMalloc(&seg_off, nt_cap * 4));
    B2RT_CUDA_OK(cudaMalloc(&cursor, nt_cap * 4));
    cudaFree(sched_scratch); sched_scratch = nullptr;
    B2RT_CUDA_OK(cudaMalloc(&sched_scratch, (4 + 3 * (nt_cap / 1024 + 2)) * 4));
    B2RT_CUDA_OK(cudaMemset(sched_scratch, 0, (4 + 3 * (nt_cap / 1024 + 2)) * 4));
  }
  if (chunk_alloc < chunk_cap) {
    cudaFree(chunks); chunks = nullptr;
    chunk_alloc = chunk_cap + chunk_cap / 4;
    B2RT_CUDA_OK(cudaMalloc(&chunks, chunk_alloc * sizeof(uint4)));
  }
  B2RT_CUDA_OK(cudaMemset(cnt, 0, nt_cap * 4));
  B2RT_CUDA_OK(cudaDeviceSynchronize());   // legacy-stream memsets vs the non-blocking work stream
  return B2RT_OK;
}

double Tracer::harvest_traverse_ms(double* ms_l0) {
  double total = 0, l0 = 0;
  for (size_t i = 0; i + 1 < ev_used; i += 2) {
    float ms = 0;
    if (cudaEventElapsedTime(&ms, ev_pool[i], ev_pool[i + 1]) == cudaSuccess) {
      total += ms;
      if (i / 2 < ev_deeper.size() && !ev_deeper[i / 2]) l0 += ms;
    }
  }
  ev_used = 0;
  ev_deeper.clear();
  if (ms_l0) *ms_l0 = l0;
  return total;
}

int Tracer::read_counters(TraceCounters* total, TraceCounters* l0) {
  TraceCounters c[2];
  memset(c, 0, sizeof c);
  if (counters) B2RT_CUDA_OK(cudaMemcpy(c, counters, sizeof c, cudaMemcpyDeviceToHost));
  if (l0) *l0 = c[0];
  if (total) {
    total->node_visits = c[0].node_visits + c[1].node_visits; total->prim_tests = c[0].prim_tests + c[1].prim_tests;
    total->subtree_visits = c[0].subtree_visits + c[1].subtree_visits; total->pushes = c[0].pushes + c[1].pushes;
    total->staged_bytes = c[0].staged_bytes + c[1].staged_bytes; total->hit_updates = c[0].hit_updates + c[1].hit_updates;
  }
  return B2RT_OK;
}

int Tracer::reset_counters(cudaStream_t s) {
  if (counters) B2RT_CUDA_OK(cudaMemsetAsync(counters, 0, 2 * sizeof(TraceCounters), s));
  return B2RT_OK;
}

void Tracer::release() {
  cudaFree(cnt); cudaFree(seg_off); cudaFree(cursor); cudaFree(pairs); cudaFree(ids_sorted); cudaFree(chunks); cudaFree(sched_scratch); sched_scratch = nullptr;
  cudaFree(ctrl); cudaFree(counters);
  cnt = seg_off = cursor = ids_sorted = ctrl = nullptr; pairs = nullptr; chunks = nullptr; counters = nullptr;
  pair_cap = 0; max_rays = 0; nt_cap = 0; chunk_alloc = 0;
  for (int k = 0; k < 2; ++k) {
    cudaFree(sl_o[k]); cudaFree(sl_d[k]); cudaFree(sl_h[k]); cudaFree(sl_map[k]); cudaFree(sl_exit[k]);
    sl_o[k] = sl_d[k] = nullptr; sl_h[k] = nullptr; sl_map[k] = nullptr; sl_exit[k] = nullptr;
  }
  cudaFree(sl_n); sl_n = nullptr; slice_cap = 0;
}

template <int W>
static void launch_traverse(const Tracer& T, cudaStream_t s, const TravParams& P, bool any_hit, bool stats) {
  dim3 grid(T.num_sms * T.ctas_per_sm), block(TRAV_THREADS);
  if (any_hit) {
    if (stats) k_traverse<W, true, true><<<grid, block, T.smem_bytes, s>>>(P);
    else k_traverse<W, true, false><<<grid, block, T.smem_bytes, s>>>(P);
  } else {
    if (stats) k_traverse<W, false, true><<<grid, block, T.smem_bytes, s>>>(P);
    else k_traverse<W, false, false><<<grid, block, T.smem_bytes, s>>>(P);
  }
}

int Tracer::trace(cudaStream_t s, const float4* ray_o, const float4* ray_d, unsigned long long* hits,
                  const uint32_t* n_active_dev, bool any_hit) {
  if (bvh.n_levels == 0) return B2RT_OK;
  // control words: pair counters, chunk cursors (level 0 uses PAIRS0 / NEXT0; higher levels are reset by their
  // scheduling kernel).  Per-subtree counts are left at zero by the previous trace (self-cleaning scheduler).
  B2RT_CUDA_OK(cudaMemsetAsync(ctrl, 0, 4 * 4, s));
  for (uint32_t L = 0; L < bvh.n_levels; ++L) {
    const LevelRange lr = bvh.levels[L];
    if (L > 0) {
      if (lr.count <= 12288) {
        k_count_tiled<<<num_sms * count_ctas, 256, (size_t)lr.count * 4, s>>>(pairs, &ctrl[(L - 1) & 1], cnt, (uint32_t)pair_cap, lr.first, lr.count);
      } else {
        k_count<<<num_sms * 8, 256, 0, s>>>(pairs, &ctrl[(L - 1) & 1], cnt, (uint32_t)pair_cap);
      }
      k_schedule_level<<<std::max(1u, (lr.count + 1023) / 1024), 1024, 0, s>>>(cnt, seg_off, cursor, chunks, ctrl, lr.first, lr.count, chunk_rays,
                                          (uint32_t)chunk_cap, L, (uint32_t)pair_cap, num_sms * ctas_per_sm * chunks_per_cta, chunk_min,
                                          sched_scratch);
      launches += 2;
    }
    if (L > 0) {
      if (lr.count <= 12288) {
        k_scatter_tiled<<<num_sms * scatter_ctas, 256, (size_t)lr.count * 8, s>>>(pairs, &ctrl[(L - 1) & 1], seg_off, cursor, ids_sorted,
                                                                        (uint32_t)pair_cap, lr.first, lr.count);
      } else {
        k_scatter<<<num_sms * 8, 256, 0, s>>>(pairs, &ctrl[(L - 1) & 1], seg_off, cursor, ids_sorted, (uint32_t)pair_cap);
      }
      launches++;
    }
    TravParams P;
    P.blob = bvh.blob; P.treelets = bvh.treelets; P.ray_o = ray_o; P.ray_d = ray_d; P.hits = hits;
    P.ids = (L == 0) ? nullptr : ids_sorted;
    P.chunks = chunks; P.ctrl = ctrl; P.cnt = cnt; P.pairs = pairs; P.pair_cap = (uint32_t)pair_cap; P.level = L;
    P.counters = counters + (L ? 1 : 0); P.n_treelets = bvh.n_treelets; P.chunk_rays = chunk_rays; P.chunk0_max = std::max(chunk0_max, chunk_rays); P.stack_off = (uint32_t)stack_off; P.ring_off = (uint32_t)ring_off; P.n_active = n_active_dev; P.n_rays_cap = (uint32_t)std::min<uint64_t>(max_rays, 0xFFFFFFFFull);
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (time_kernels) {
      if (ev_used + 2 > ev_pool.size()) {
        for (int k = 0; k < 64; ++k) { cudaEvent_t e; B2RT_CUDA_OK(cudaEventCreate(&e)); ev_pool.push_back(e); }
      }
      e0 = ev_pool[ev_used++]; e1 = ev_pool[ev_used++];
      ev_deeper.push_back(L ? 1 : 0);
      cudaEventRecord(e0, s);
    }
    if (bvh.width == 2) launch_traverse<2>(*this, s, P, any_hit, collect_stats);
    else if (bvh.width == 8) launch_traverse<8>(*this, s, P, any_hit, collect_stats);
    else if (bvh.width == 16) launch_traverse<16>(*this, s, P, any_hit, collect_stats);
    else launch_traverse<4>(*this, s, P, any_hit, collect_stats);
    if (time_kernels) cudaEventRecord(e1, s);
    launches++; traverse_launches++; if (L == 0) traverse_launches_l0++;
  }
  B2RT_CUDA_OK(cudaGetLastError());
  return B2RT_OK;
}

int Tracer::ensure_slices(uint64_t n) {
  if (slice_cap >= n && sl_n) return B2RT_OK;
  for (int k = 0; k < 2; ++k) {
    cudaFree(sl_o[k]); cudaFree(sl_d[k]); cudaFree(sl_h[k]); cudaFree(sl_map[k]); cudaFree(sl_exit[k]);
    sl_o[k] = sl_d[k] = nullptr; sl_h[k] = nullptr; sl_map[k] = nullptr; sl_exit[k] = nullptr;
  }
  slice_cap = 0;
  for (int k = 0; k < 2; ++k) {
    B2RT_CUDA_OK(cudaMalloc(&sl_o[k], (n + 16) * sizeof(float4)));
    B2RT_CUDA_OK(cudaMalloc(&sl_d[k], (n + 16) * sizeof(float4)));
    B2RT_CUDA_OK(cudaMalloc(&sl_h[k], (n + 16) * 8));
    B2RT_CUDA_OK(cudaMalloc(&sl_map[k], (n + 16) * 4));
    B2RT_CUDA_OK(cudaMalloc(&sl_exit[k], (n + 16) * 4));
  }
  if (!sl_n) B2RT_CUDA_OK(cudaMalloc(&sl_n, 2 * 4));
  slice_cap = n;
  return B2RT_OK;
}

int Tracer::trace_sliced(cudaStream_t s, const float4* ray_o, const float4* ray_d, unsigned long long* hits,
                         const uint32_t* n_active_dev, uint64_t n_max, bool any_hit) {
  if (!(slice_first > 0.f) || slice_passes < 2 || bvh.n_levels == 0) return trace(s, ray_o, ray_d, hits, n_active_dev, any_hit);
  if (n_max > max_rays) { set_error("trace_sliced: batch larger than the tracer was initialised for"); return B2RT_ERR_INVALID; }
  int rc = ensure_slices(n_max);
  if (rc) return rc;
  SliceBox box;
  for (int k = 0; k < 6; ++k) box.v[k] = slice_bbox[k];
  const float ex = slice_bbox[3] - slice_bbox[0], ey = slice_bbox[4] - slice_bbox[1], ez = slice_bbox[5] - slice_bbox[2];
  box.margin = 1e-3f * std::sqrt(ex * ex + ey * ey + ez * ez);
  SliceList L[2];
  for (int k = 0; k < 2; ++k) L[k] = SliceList{sl_o[k], sl_d[k], sl_h[k], sl_map[k], sl_exit[k], sl_n + k};
  const dim3 grid(num_sms * 8);
  B2RT_CUDA_OK(cudaMemsetAsync(sl_n, 0, 8, s));
  k_slice_init<<<grid, 256, 0, s>>>(n_active_dev, ray_o, ray_d, hits, box, slice_first, any_hit, L[0]);
  launches++;
  float len = slice_first;
  for (int p = 0; p < slice_passes; ++p) {
    const int cur = p & 1, nxt = cur ^ 1;
    const bool last = p + 1 == slice_passes;
    rc = trace(s, sl_o[cur], sl_d[cur], sl_h[cur], sl_n + cur, any_hit);
    if (rc) return rc;
    len *= slice_growth;
    B2RT_CUDA_OK(cudaMemsetAsync(sl_n + nxt, 0, 4, s));
    // the pass before the last hands the whole remainder to the last one
    k_slice_advance<<<grid, 256, 0, s>>>(L[cur], ray_d, hits, p + 2 == slice_passes ? __builtin_huge_valf() : len, last, L[nxt]);
    launches++;
  }
  B2RT_CUDA_OK(cudaGetLastError());
  return B2RT_OK;
}

int Tracer::check_overflow(cudaStream_t s, bool* overflow) {
  uint32_t v = 0;
  B2RT_CUDA_OK(cudaMemcpyAsync(&v, ctrl + CTRL_OVERFLOW, 4, cudaMemcpyDeviceToHost, s));
  B2RT_CUDA_OK(cudaStreamSynchronize(s));
  *overflow = v != 0;
#ifdef B2RT_CHECKS
  uint32_t dbg[2] = {0, 0};
  cudaMemcpy(dbg, ctrl + 6, 8, cudaMemcpyDeviceToHost);
  if (dbg[0]) {
    fprintf(stderr, "B2RT_CHECKS: code %u info %u (0x%08x)\n", dbg[0], dbg[1], dbg[1]);
    cudaMemset(ctrl + 6, 0, 8);
    set_error("B2RT_CHECKS: device invariant " + std::to_string(dbg[0]) + " violated (info " + std::to_string(dbg[1]) + ")");
    return B2RT_ERR_CUDA;
  }
#endif
  if (v) B2RT_CUDA_OK(cudaMemsetAsync(ctrl + CTRL_OVERFLOW, 0, 4, s));
  return B2RT_OK;
}

}  // namespace b2rt
