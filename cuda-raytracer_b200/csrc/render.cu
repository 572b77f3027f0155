// Wavefront path tracer around the traversal path: ray generation, per-bounce shading (emission,
// next-event estimation, BSDF sampling), shadow resolve, accumulation, 3x3 median, tone map.
//
// Replaces (fused / re-designed) the reference kernels
//   kernelPrimaryRays           src/cudaRenderer.cu:312-376   -> k_raygen (Philox per (pixel,sample))
//   kernelDirectLightRays       src/cudaRenderer.cu:380-481   -> k_shade (NEE part) + k_resolve_shadow
//   kernelProcessIntersections  src/cudaRenderer.cu:544-664   -> k_shade (BSDF part)
//   kernelUpdateSSImage / kernelReconstructImage / kernelAccumulate  :666-747 -> k_accumulate
//   kernelMedianFilter          src/cudaRenderer.cu:773-842   -> k_median3x3 (shared-memory tile)
// and the host driver renderFrame / renderAccumulate (:2419-2564) without any per-kernel
// cudaDeviceSynchronize or per-level host read-back.  The estimator is the Scotty3D one
// (src/pathtracer.cpp:395-496 skeleton) with the runtime knobs ns_aa / max_ray_depth /
// ns_area_light; see DESIGN.md "Integrator" for the exact definition shared with the oracle.
#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>
#include <cmath>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "render.cuh"
#include "rt_device.cuh"

namespace b2rt {

namespace {

constexpr float INF_F = __builtin_huge_valf();
constexpr uint32_t MAX_DEPTH = 96, ACT0 = 8, SH0 = ACT0 + MAX_DEPTH + 8, SHV0 = SH0 + MAX_DEPTH + 8, REAL0 = SHV0 + MAX_DEPTH + 8,
                   BLKA0 = REAL0 + MAX_DEPTH + 8, BLKS0 = BLKA0 + MAX_DEPTH + 8, N_COUNTS = 1024;
static_assert(BLKS0 + MAX_DEPTH + 8 <= N_COUNTS, "counter table");

struct CamDev { f3 pos, cx, cy, cz; float tan_h, tan_v; };

struct SceneDev {
  const float4* prim_geom;      // 3 float4 per prim, scene order
  const float* tri_normals;     // 9 per tri or nullptr
  const uint32_t* prim_material;
  const b2rt_material* materials;
  const b2rt_light* lights;
  const float* light_area;
  uint32_t n_tris, n_lights;
  const float* env;             // environment map, RGB triples, index x + y*w (row 0 = the +y pole), or nullptr
  uint32_t env_w, env_h;
};

// Rays live in DENSE lists (what level 0 of the traversal streams): entry i of the bounce-b list is the ray of path
// `lslot[i]`; shading appends the continuing paths to the other list in warp-private blocks, whose unfilled tails are
// null entries (slot 0xFFFFFFFF, empty interval).  Per-path state (throughput, radiance) is indexed by the path's slot.
struct PathBufs {
  // current bounce's list (read) and the next bounce's (appended to); the host swaps them per bounce -- plain
  // pointers, not an indexed array: a dynamically indexed kernel parameter would be copied to local memory
  float4* lo; float4* ld; unsigned long long* lh; uint32_t* lslot;
  float4* no; float4* nd; unsigned long long* nh; uint32_t* nslot;
  float4* thr;     // [slot] rgb throughput, w = count_emission flag
  float4* rad;     // [slot] rgb radiance of the path
  // dense shadow-ray list of the current bounce: a path that hit a diffuse surface owns S consecutive entries
  // (one per light sample; samples that cannot contribute are null rays the any-hit traversal skips)
  float4* s_o; float4* s_d; unsigned long long* s_hits; float4* s_contrib;
  uint32_t* s_q0;    // [slot] first shadow-list entry of the path's block, 0xFFFFFFFF = none
  // [3] = cancel flag; [ACT0 + b] = entries of bounce b's path list (of which [REAL0 + b] are paths, the rest null
  // entries, see k_shade); [SH0 + b] = shadow-list entries of bounce b; [SHV0 + b] = of which real shadow rays;
  // [BLKA0 + b] / [BLKS0 + b] = reservation blocks handed out of the two lists
  uint32_t* counts;
  uint32_t inc_bound;   // 0xFFFFFFFF (see next_block)
  uint32_t list_cap, shadow_cap;   // entries allocated per path list / for the shadow list (checked in B2RT_CHECKS builds)
};

__global__ void __launch_bounds__(256)
k_raygen(WaveParams wp, CamDev cam, PathBufs pb) {
  const uint32_t n = wp.n_pix * wp.spp;
  const uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x;
  if (blockIdx.x == 0 && threadIdx.x < MAX_DEPTH + 1) {   // per-bounce list counters of this wave
    pb.counts[ACT0 + threadIdx.x] = (threadIdx.x == 0) ? (pb.counts[3] ? 0u : n) : 0u;
    pb.counts[SH0 + threadIdx.x] = 0u;
    pb.counts[SHV0 + threadIdx.x] = 0u;
    pb.counts[REAL0 + threadIdx.x] = 0u;
    pb.counts[BLKA0 + threadIdx.x] = 0u;
    pb.counts[BLKS0 + threadIdx.x] = 0u;
  }
  if (slot >= n) return;
  const uint32_t pix = wp.pix0 + slot / wp.spp;
  const uint32_t sample = __ldg(wp.sample_base) + wp.sample0 + (slot % wp.spp) * wp.sample_stride;
  const uint32_t x = pix % wp.width, y = pix / wp.width;
  const uint4 r = philox4x32_10(pix, sample, 0, 0, wp.k0, wp.k1);
  float jx = 0.5f, jy = 0.5f;
  if (wp.jitter) { jx = u01(r.x); jy = u01(r.y); }
  const float sx = ((float)x + jx) / (float)wp.width, sy = ((float)y + jy) / (float)wp.height;
  // Camera::generate_ray contract, src/camera.h:71-81
  const float px = (2.0f * sx - 1.0f) * cam.tan_h;
  const float py = (2.0f * sy - 1.0f) * cam.tan_v;
  const f3 w = cam.cx * px + cam.cy * py - cam.cz;
  const f3 d = normalize3(w);
  pb.lo[slot] = make_float4(cam.pos.x, cam.pos.y, cam.pos.z, 0.0f);
  pb.ld[slot] = make_float4(d.x, d.y, d.z, INF_F);
  pb.lh[slot] = pack_hit(INF_F, 0xFFFFFFFFu);
  pb.lslot[slot] = slot;
  pb.thr[slot] = make_float4(1.f, 1.f, 1.f, 1.f);
  pb.rad[slot] = make_float4(0.f, 0.f, 0.f, 0.f);
}

// bilinear look-up of the environment map in direction d (unit): wraps in azimuth, clamps at the poles.  Not inlined:
// only scenes with an environment map pay for its registers.
__device__ __noinline__ f3 env_lookup(const float* __restrict__ env, uint32_t ew, uint32_t eh, f3 d) {
  const float u = atan2_turns(d.z, d.x);
  const float cy = d.y > 1.0f ? 1.0f : (d.y < -1.0f ? -1.0f : d.y);
  const float sy = 1.0f - cy * cy;
  const float v = 2.0f * atan2_turns(__fsqrt_rn(sy < 0.0f ? 0.0f : sy), cy);
  const float fx = u * (float)ew - 0.5f, fy = v * (float)eh - 0.5f;
  const float flx = floorf(fx), fly = floorf(fy);
  const float tx = fx - flx, ty = fy - fly;
  int x0 = (int)flx, y0 = (int)fly;
  int x1 = x0 + 1, y1 = y0 + 1;
  const int W = (int)ew, H = (int)eh;
  x0 = ((x0 % W) + W) % W; x1 = ((x1 % W) + W) % W;
  y0 = y0 < 0 ? 0 : (y0 > H - 1 ? H - 1 : y0); y1 = y1 < 0 ? 0 : (y1 > H - 1 ? H - 1 : y1);
  auto px = [&](int x, int y) { const float* p = env + 3 * ((size_t)y * W + x); return mk3(__ldg(p), __ldg(p + 1), __ldg(p + 2)); };
  const f3 a = px(x0, y0) * (1.0f - tx) + px(x1, y0) * tx;
  const f3 b = px(x0, y1) * (1.0f - tx) + px(x1, y1) * tx;
  return a * (1.0f - ty) + b * ty;
}

// f(wo, wi) of the glossy BSDF in the local frame: reflectance * (n + 2) / (2 pi) * cos^n(wi, mirror direction of wo)
__device__ __noinline__ f3 glossy_f(const b2rt_material* mp, f3 wo, f3 wi) {
  const uint32_t n = glossy_exponent(__ldg(&mp->roughness));
  float c = dot3(mk3(-wo.x, -wo.y, wo.z), wi);
  if (!(c > 0.0f)) c = 0.0f;
  return mk3(__ldg(&mp->albedo[0]), __ldg(&mp->albedo[1]), __ldg(&mp->albedo[2])) * (((float)(n + 2u) * 0.159154943091895336f) * powi(c, n));
}

// make_coord_space, src/bsdf.cpp:14-33
__device__ __forceinline__ void make_coord_space(f3 n, f3* X, f3* Y, f3* Z) {
  f3 z = n, h = z;
  if (fabsf(h.x) <= fabsf(h.y) && fabsf(h.x) <= fabsf(h.z)) h.x = 1.0f;
  else if (fabsf(h.y) <= fabsf(h.x) && fabsf(h.y) <= fabsf(h.z)) h.y = 1.0f;
  else h.z = 1.0f;
  z = normalize3(z);
  f3 y = normalize3(cross3(h, z));
  f3 x = normalize3(cross3(z, y));
  *X = x; *Y = y; *Z = z;
}

// A warp's private piece of a dense output list.  Warps reserve blocks of `blk` entries from the list's block counter
// (one atomic per block instead of one per tile; the next block is reserved a tile ahead, so nobody waits for the
// atomic's round trip) and append into their piece with ballots only -- no shared memory, no barrier.  What a warp has
// reserved but not filled when it runs out of input becomes NULL entries (see k_shade).
struct OutRange {
  uint32_t next = 0, end = 0;   // warp-uniform: unused entries [next, end) of the current block
  uint32_t spare = 0;           // lane 0 only: index of the block reserved ahead
  uint32_t blocks = 0;          // warp-uniform: blocks this warp has reserved
  bool has = false;             // warp-uniform: a block is reserved ahead
};
// The block counter is bumped with atom.inc and a bound (0xFFFFFFFF) that arrives as a kernel argument.  A plain
// atom.add here is rewritten by ptxas into its warp-aggregated form -- vote, leader, ATOMG, SHFL of the result to every
// participant -- and that SHFL sits right behind the atomic and waits for the round trip this scheme exists to hide
// (13 % of the kernel's stall samples); an inc with a run-time bound is left alone.
__device__ __forceinline__ uint32_t next_block(uint32_t* block_counter, uint32_t bound) {
  uint32_t old;
  asm volatile("atom.global.inc.u32 %0, [%1], %2;" : "=r"(old) : "l"(block_counter), "r"(bound) : "memory");
  return old;
}
__device__ __forceinline__ void reserve_ahead(OutRange& r, uint32_t* block_counter, uint32_t bound, uint32_t need, uint32_t lane) {
  if (!r.has && r.end - r.next < need) {
    if (lane == 0) r.spare = next_block(block_counter, bound);
    r.has = true; r.blocks++;
  }
}
// position of this lane's run of `per` entries (garbage when pred == false); blk_entries >= 32 * per
__device__ __forceinline__ uint32_t warp_append(OutRange& r, uint32_t* block_counter, uint32_t bound, bool pred, uint32_t per,
                                                uint32_t blk_entries, uint32_t lane, uint32_t lane_lt) {
  const uint32_t m = __ballot_sync(0xffffffffu, pred);
  const uint32_t cnt = (uint32_t)__popc(m) * per, rank = (uint32_t)__popc(m & lane_lt) * per;
  const uint32_t avail = r.end - r.next;
  uint32_t pos;
  if (cnt > avail) {   // warp-uniform: the block is full, the rest goes to the head of the next one
    if (!r.has) { if (lane == 0) r.spare = next_block(block_counter, bound); r.blocks++; }
    const uint32_t base = __shfl_sync(0xffffffffu, r.spare, 0) * blk_entries;
    pos = rank < avail ? r.next + rank : base + (rank - avail);
    r.next = base + (cnt - avail); r.end = base + blk_entries; r.has = false;
  } else {
    pos = r.next + rank;
    r.next += cnt;
  }
  return pos;
}

// One surface interaction for every active path (bounce index b).
// PERSISTENT warps: the launch has a fixed number of warps; warp w takes the 32-entry tiles w, w + stride, ... of the
// bounce's path list, loads the next tile's list entries before it works on the current one, and appends the continuing
// paths and the shadow rays to warp-private pieces of the output lists (OutRange).  Round 1 ran one thread per entry
// with two CTA-aggregated appends; their barriers -- every warp of the CTA waiting for the one global atomic behind
// the slowest warp's dependent loads -- were 35 % of the kernel's warp samples (profiles/r01_shade_ncu.md).
// NULL entries: the unfilled rest of a warp's last block.  Path list: slot 0xFFFFFFFF, hit word (t = -1, prim 0);
// shadow list: contribution 0, hit word (t = -1, prim 0).  The traversal retires a ray whose interval is empty without
// a node visit, k_shade / k_resolve_shadow skip entries without a slot.  A list is at most
// shade_slack() entries longer than the paths it holds.
#ifndef B2RT_SHADE_OCC
#define B2RT_SHADE_OCC 5
#endif
#ifndef B2RT_SHADE_THREADS
#define B2RT_SHADE_THREADS 128
#endif
constexpr int SHADE_THREADS = B2RT_SHADE_THREADS;
constexpr uint32_t SHADE_BLK_MAX = 128;   // largest reservation block, in list items
constexpr unsigned long long NULL_HIT = 0xBF80000000000000ull;   // (t = -1.0f, prim 0): an empty interval, "occluded"
// EXT = the scene has an environment map or a glossy material; the common instantiation carries neither (their
// look-up / lobe calls cost registers the 64-register kernel does not have to spare)
template <bool EXT>
__global__ void __launch_bounds__(SHADE_THREADS, B2RT_SHADE_OCC)
k_shade(WaveParams wp, SceneDev sc, PathBufs pb, uint32_t b) {
  const uint32_t n = pb.counts[ACT0 + b];
  const uint32_t lane = threadIdx.x & 31, lane_lt = (1u << lane) - 1u;
  const uint32_t n_tiles = (n + 31u) >> 5;
  // short lists use fewer warps (>= 8 tiles each), so that the null entries stay a small part of the output
  const uint32_t stride = min((gridDim.x * (uint32_t)SHADE_THREADS) >> 5, max(1u, n_tiles >> 3));
  const uint32_t gw = (blockIdx.x * (uint32_t)SHADE_THREADS + threadIdx.x) >> 5;
  if (gw >= stride) return;
  const uint32_t tiles_per_warp = n_tiles / stride;
  const uint32_t blk = tiles_per_warp < 32u ? 32u : (tiles_per_warp < 128u ? 64u : SHADE_BLK_MAX);
  const uint32_t S = wp.S;
  const bool single = S == 1;
  const uint32_t blk_sh = blk * max(S, 1u);
  OutRange out_a, out_s;
  uint32_t real_a = 0, real_s = 0;   // per-lane counts of the paths / real shadow rays this lane appended
  // the next tile's list entries, loaded one tile ahead
  uint32_t p_slot = 0xFFFFFFFFu;
  unsigned long long p_h = 0;
  float4 p_o = make_float4(0.f, 0.f, 0.f, 0.f), p_d = make_float4(0.f, 0.f, 1.f, 0.f);
  {
    const uint32_t i = gw * 32u + lane;
    if (i < n) { p_slot = pb.lslot[i]; p_h = pb.lh[i]; p_o = pb.lo[i]; p_d = pb.ld[i]; }
  }
  for (uint32_t tile = gw; tile < n_tiles; tile += stride) {
  const uint32_t slot = p_slot;
  const unsigned long long h = p_h;
  const float4 ro = p_o, rd = p_d;
  const bool active = slot != 0xFFFFFFFFu;   // not past the end of the list, not a null entry
  p_slot = 0xFFFFFFFFu;
  {
    const uint32_t i = (tile + stride) * 32u + lane;
    if (tile + stride < n_tiles && i < n) { p_slot = pb.lslot[i]; p_h = pb.lh[i]; p_o = pb.lo[i]; p_d = pb.ld[i]; }
  }
  reserve_ahead(out_a, &pb.counts[BLKA0 + b + 1], pb.inc_bound, 32u, lane);
  if (S > 0) reserve_ahead(out_s, &pb.counts[BLKS0 + b], pb.inc_bound, 32u * S, lane);
  bool cont = false;
  float4 next_o = make_float4(0.f, 0.f, 0.f, 0.f), next_d = make_float4(0.f, 0.f, 1.f, 0.f);
  uint32_t prim = 0xFFFFFFFFu;
  // the material is read field by field where it is used (the table is a few cache lines); holding the 48-byte record
  // and the 64-byte light record in registers across the kernel was the main source of spills at 64 registers
  const b2rt_material* mp = sc.materials;
  int32_t m_kind = -1;
  float4 thr4 = make_float4(0.f, 0.f, 0.f, 0.f);
  PrimRec pr;
  pr.a = pr.b = pr.c = make_float4(0.f, 0.f, 0.f, 0.f);
  float vn[9] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (active) {
    thr4 = pb.thr[slot];
    prim = (uint32_t)h;
    if (prim != 0xFFFFFFFFu) {
      // everything that depends on the primitive index alone is requested at once: the material word (index | kind << 28,
      // k_tag_material), the primitive record, the vertex normals
      const uint32_t pm = __ldg(sc.prim_material + prim);
      pr.a = __ldg(sc.prim_geom + (size_t)prim * 3); pr.b = __ldg(sc.prim_geom + (size_t)prim * 3 + 1); pr.c = __ldg(sc.prim_geom + (size_t)prim * 3 + 2);
      if (sc.tri_normals && prim < sc.n_tris) {
        const float* nn = sc.tri_normals + (size_t)prim * 9;
#pragma unroll
        for (int k = 0; k < 9; ++k) vn[k] = __ldg(nn + k);
      }
      mp = sc.materials + (pm & 0x0FFFFFFFu); m_kind = (int32_t)(pm >> 28);
    }
  }
  // every diffuse hit reserves S consecutive entries of the bounce's shadow-ray list: one atomic per warp, blocks in
  // lane order, so the list stays coalesced against the path list
  // S > 1: every diffuse hit reserves S consecutive entries up front (samples that cannot contribute become null rays, so
  // the block keeps its sample order for the deterministic resolve).  S == 1 (one light sample per interaction, every
  // BASELINE config): the one sample is kept in registers and appended at the end ONLY if it can contribute -- on the
  // box scenes 45 % of the light samples lie behind the surface, and null rays cost the any-hit traversal a list
  // entry each (stream + retire).
  const bool wants_shadow = prim != 0xFFFFFFFFu && (m_kind == B2RT_MAT_DIFFUSE || (EXT && m_kind == B2RT_MAT_GLOSSY)) && S > 0;
  uint32_t q0 = 0xFFFFFFFFu;
  if (!single && S > 0) {
    q0 = warp_append(out_s, &pb.counts[BLKS0 + b], pb.inc_bound, wants_shadow, S, blk_sh, lane, lane_lt);
    if (!wants_shadow) q0 = 0xFFFFFFFFu;
  }
  uint32_t n_valid = 0;
  bool sh_valid = false;
  float4 sh_d = make_float4(0.f, 0.f, 1.f, -1.f);
  f3 sh_c = mk3(0.f, 0.f, 0.f);
  if (active) {
    if (!single) pb.s_q0[slot] = q0;
    if (prim == 0xFFFFFFFFu) {
      // the ray left the scene: EnvironmentLight::sample_dir, counted like emitted radiance (camera rays, delta bounces)
      if (EXT && sc.env && thr4.w != 0.f) {
        const f3 e = env_lookup(sc.env, sc.env_w, sc.env_h, mk3(rd.x, rd.y, rd.z));
        float4 L = pb.rad[slot];
        const f3 add = mk3(thr4.x, thr4.y, thr4.z) * e;
        L.x = L.x + add.x; L.y = L.y + add.y; L.z = L.z + add.z;
        pb.rad[slot] = L;
      }
    } else {
      const float t = __uint_as_float((uint32_t)(h >> 32));
      f3 thr = mk3(thr4.x, thr4.y, thr4.z);
      const bool count_emission = thr4.w != 0.f;
      if (m_kind == B2RT_MAT_EMISSION) {
        if (count_emission) {
          float4 L = pb.rad[slot];
          const f3 add = thr * mk3(__ldg(&mp->emission[0]), __ldg(&mp->emission[1]), __ldg(&mp->emission[2]));
          L.x = L.x + add.x; L.y = L.y + add.y; L.z = L.z + add.z;
          pb.rad[slot] = L;
        }
      } else {
        const f3 o = mk3(ro.x, ro.y, ro.z), d = mk3(rd.x, rd.y, rd.z);
        const f3 P = o + d * t;
        next_o = make_float4(P.x, P.y, P.z, wp.eps);   // origin of the shadow ray and of the continuing ray
        const bool is_sphere = prim >= sc.n_tris;
        f3 nrm;
        bool backface = false;
        if (is_sphere) {
          nrm = normalize3(P - mk3(pr.a.x, pr.a.y, pr.a.z));
        } else if (sc.tri_normals) {
          // recover (u,v) with the same arithmetic as the traversal test
          float tt, u = 0.f, v = 0.f;
          hit_triangle(pr, o, d, ro.w, INF_F, &tt, &u, &v);
          const float w0 = 1.0f - u - v;
          nrm = mk3(vn[3], vn[4], vn[5]) * u + mk3(vn[6], vn[7], vn[8]) * v + mk3(vn[0], vn[1], vn[2]) * w0;
        } else {
          nrm = cross3(mk3(pr.a.w, pr.b.x, pr.b.y), mk3(pr.b.z, pr.b.w, pr.c.x));
        }
        if (!(dot3(d, nrm) < 0.0f)) { nrm = neg3(nrm); backface = true; }
        if (!is_sphere) nrm = normalize3(nrm);
        f3 X, Y, Z;
        make_coord_space(nrm, &X, &Y, &Z);
        const f3 wo_w = neg3(d);
        const f3 wo = mk3(dot3(wo_w, X), dot3(wo_w, Y), dot3(wo_w, Z));
        const uint32_t pix = wp.pix0 + slot / wp.spp;
        const uint32_t sample = __ldg(wp.sample_base) + wp.sample0 + (slot % wp.spp) * wp.sample_stride;

        if (m_kind == B2RT_MAT_DIFFUSE || (EXT && m_kind == B2RT_MAT_GLOSSY)) {
          // direct lighting: src/pathtracer.cpp:439-478 + shadow ray (Task 4); the environment map is one more light
          uint32_t j = 0;
          const uint32_t n_lights_all = sc.n_lights + ((EXT && sc.env) ? 1u : 0u);
          for (uint32_t li = 0; li < n_lights_all; ++li) {
            const bool is_env = EXT && li == sc.n_lights;
            const b2rt_light* lt = sc.lights + (is_env ? 0u : li);     // (never dereferenced for the environment light)
            const int32_t lt_kind = is_env ? 3 : __ldg(&lt->kind);
            const uint32_t ns = lt_kind == B2RT_LIGHT_AREA ? max(1u, wp.ns_area_light) : 1u;
            for (uint32_t k = 0; k < ns; ++k, ++j) {
              f3 wi; float dist, pdf; f3 Lr;
              f3 lp = mk3(0, 0, 0), ld = mk3(0, 0, 0), radc = mk3(0, 0, 0);
              if (!is_env) {
                lp = mk3(__ldg(&lt->position[0]), __ldg(&lt->position[1]), __ldg(&lt->position[2]));
                ld = mk3(__ldg(&lt->direction[0]), __ldg(&lt->direction[1]), __ldg(&lt->direction[2]));
                radc = mk3(__ldg(&lt->radiance[0]), __ldg(&lt->radiance[1]), __ldg(&lt->radiance[2]));
              }
              if (is_env) {
                // EnvironmentLight::sample_L, uniform over the sphere: pdf = 1 / (4 pi)
                const uint4 r4 = philox4x32_10(pix, sample, b, 1 + j, wp.k0, wp.k1);
                const float zz = 1.0f - 2.0f * u01(r4.x);
                float sn, cs;
                sincos2pi(u01(r4.y), &sn, &cs);
                const float rr2 = 1.0f - zz * zz;
                const float rr = __fsqrt_rn(rr2 < 0.0f ? 0.0f : rr2);
                wi = mk3(rr * cs, zz, rr * sn);
                dist = INF_F; pdf = 0.0795774715459476679f;
                Lr = env_lookup(sc.env, sc.env_w, sc.env_h, wi);
              } else if (lt_kind == B2RT_LIGHT_AREA) {
                // AreaLight::sample_L, src/static_scene/light.cpp:81-92
                const uint4 r4 = philox4x32_10(pix, sample, b, 1 + j, wp.k0, wp.k1);
                const float ux = u01(r4.x) - 0.5f, uy = u01(r4.y) - 0.5f;
                const f3 dv = lp + mk3(__ldg(&lt->dim_x[0]), __ldg(&lt->dim_x[1]), __ldg(&lt->dim_x[2])) * ux + mk3(__ldg(&lt->dim_y[0]), __ldg(&lt->dim_y[1]), __ldg(&lt->dim_y[2])) * uy - P;
                const float sq = dot3(dv, dv);
                dist = __fsqrt_rn(sq);
                const float invd = __fdiv_rn(1.0f, dist);
                wi = dv * invd;
                const float cosT = dot3(wi, ld);
                pdf = __fdiv_rn(sq, sc.light_area[li] * fabsf(cosT));
                Lr = cosT < 0.0f ? radc : mk3(0, 0, 0);
              } else if (lt_kind == B2RT_LIGHT_POINT) {
                const f3 dv = lp - P;
                const float sq = dot3(dv, dv);
                dist = __fsqrt_rn(sq);
                wi = dv * __fdiv_rn(1.0f, dist);
                pdf = 1.0f; Lr = radc;
              } else {
                wi = neg3(ld); dist = INF_F; pdf = 1.0f; Lr = radc;
              }
              const float cos_in = dot3(wi, Z);
              const bool valid = cos_in >= 0.0f && (Lr.x > 0.0f || Lr.y > 0.0f || Lr.z > 0.0f) && pdf > 0.0f;
              const uint32_t q = q0 + j;
              if (single) {
                if (valid) {
                  const float wgt = __fdiv_rn(cos_in, (float)ns * pdf);
                  const f3 f = (EXT && m_kind == B2RT_MAT_GLOSSY) ? glossy_f(mp, wo, mk3(dot3(wi, X), dot3(wi, Y), cos_in))
                                                         : mk3(__ldg(&mp->albedo[0]), __ldg(&mp->albedo[1]), __ldg(&mp->albedo[2])) * 0.318309886183790672f;
                  sh_c = thr * f * Lr * wgt;
                  sh_d = make_float4(wi.x, wi.y, wi.z, dist - wp.eps);
                  sh_valid = true;
                  n_valid++;
                }
              } else if (valid) {
                const float wgt = __fdiv_rn(cos_in, (float)ns * pdf);
                const f3 f = (EXT && m_kind == B2RT_MAT_GLOSSY) ? glossy_f(mp, wo, mk3(dot3(wi, X), dot3(wi, Y), cos_in))
                                                       : mk3(__ldg(&mp->albedo[0]), __ldg(&mp->albedo[1]), __ldg(&mp->albedo[2])) * 0.318309886183790672f;
                const f3 c = thr * f * Lr * wgt;
                const float tmx = dist - wp.eps;
                pb.s_o[q] = make_float4(P.x, P.y, P.z, wp.eps);
                pb.s_d[q] = make_float4(wi.x, wi.y, wi.z, tmx);
                pb.s_hits[q] = pack_hit(tmx, 0xFFFFFFFFu);
                pb.s_contrib[q] = make_float4(c.x, c.y, c.z, 1.f);
                n_valid++;
              } else {
                // null entry: hit word already "occluded", so the any-hit traversal retires it without a node visit
                pb.s_o[q] = make_float4(0.f, 0.f, 0.f, 0.f);
                pb.s_d[q] = make_float4(0.f, 0.f, 1.f, -1.f);
                pb.s_hits[q] = pack_hit(0.f, 0u);
                pb.s_contrib[q] = make_float4(0.f, 0.f, 0.f, 0.f);
              }
            }
          }
        }
        if (b + 1 < wp.max_depth) {
          const uint4 r4 = philox4x32_10(pix, sample, b, 0, wp.k0, wp.k1);
          const float u2 = u01(r4.z), u3 = u01(r4.w);
          f3 wi_l, weight;
          bool delta = false;
          if (m_kind == B2RT_MAT_DIFFUSE || (EXT && m_kind == B2RT_MAT_GLOSSY)) {
            const float r = __fsqrt_rn(u2);
            float s, c;
            sincos2pi(u3, &s, &c);
            const float zz = 1.0f - u2;
            wi_l = mk3(r * c, r * s, __fsqrt_rn(zz < 0.0f ? 0.0f : zz));
            // cosine-weighted sample, pdf = cos / pi: f * cos / pdf = f * pi (= albedo for the diffuse BSDF)
            weight = (!EXT || m_kind == B2RT_MAT_DIFFUSE) ? mk3(__ldg(&mp->albedo[0]), __ldg(&mp->albedo[1]), __ldg(&mp->albedo[2]))
                                                          : glossy_f(mp, wo, wi_l) * 3.14159265358979324f;
          } else if (m_kind == B2RT_MAT_MIRROR) {
            wi_l = mk3(-wo.x, -wo.y, wo.z);
            weight = mk3(__ldg(&mp->albedo[0]), __ldg(&mp->albedo[1]), __ldg(&mp->albedo[2]));
            delta = true;
          } else {
            delta = true;
            const float eta = backface ? __ldg(&mp->ior) : __fdiv_rn(1.0f, __ldg(&mp->ior));
            const float cos_i = wo.z;
            const float sin2_t = eta * eta * (1.0f - cos_i * cos_i);
            const bool tir = !(sin2_t < 1.0f);
            const float cos_t = tir ? 0.0f : __fsqrt_rn(1.0f - sin2_t);
            float Fr = 1.0f;
            if (!tir) {
              const float ni = backface ? __ldg(&mp->ior) : 1.0f, nt = backface ? 1.0f : __ldg(&mp->ior);
              const float rs = __fdiv_rn(ni * cos_i - nt * cos_t, ni * cos_i + nt * cos_t);
              const float rp = __fdiv_rn(nt * cos_i - ni * cos_t, nt * cos_i + ni * cos_t);
              Fr = 0.5f * (rs * rs + rp * rp);
            }
            bool reflect;
            if (m_kind == B2RT_MAT_GLASS) reflect = tir || (u2 < Fr);
            else reflect = tir;
            if (reflect) {
              wi_l = mk3(-wo.x, -wo.y, wo.z);
              weight = m_kind == B2RT_MAT_GLASS ? mk3(__ldg(&mp->albedo[0]), __ldg(&mp->albedo[1]), __ldg(&mp->albedo[2])) : mk3(1, 1, 1);
            } else {
              wi_l = mk3(-wo.x * eta, -wo.y * eta, -cos_t);
              weight = mk3(__ldg(&mp->transmittance[0]), __ldg(&mp->transmittance[1]), __ldg(&mp->transmittance[2]));
            }
          }
          if (weight.x > 0.0f || weight.y > 0.0f || weight.z > 0.0f) {
            thr = thr * weight;
            const f3 nd = normalize3(X * wi_l.x + Y * wi_l.y + Z * wi_l.z);
            next_d = make_float4(nd.x, nd.y, nd.z, INF_F);
            pb.thr[slot] = make_float4(thr.x, thr.y, thr.z, delta ? 1.f : 0.f);
            cont = true;
          }
        }
      }
    }
  }
  // the continuing paths form the next bounce's dense ray list
  const uint32_t p = warp_append(out_a, &pb.counts[BLKA0 + b + 1], pb.inc_bound, cont, 1, blk, lane, lane_lt);
  if (single) {
    const uint32_t q = warp_append(out_s, &pb.counts[BLKS0 + b], pb.inc_bound, sh_valid, 1, blk_sh, lane, lane_lt);
    if (sh_valid) {
      // the entry names its path (w = slot bits): k_resolve_shadow1 walks the shadow list itself
      pb.s_o[q] = next_o; pb.s_d[q] = sh_d; pb.s_hits[q] = pack_hit(sh_d.w, 0xFFFFFFFFu);
      pb.s_contrib[q] = make_float4(sh_c.x, sh_c.y, sh_c.z, __uint_as_float(slot));
    }
  }
  real_s += n_valid;
#ifdef B2RT_CHECKS
  if ((cont && p >= pb.list_cap) || out_a.end > pb.list_cap || out_s.end > pb.shadow_cap) __trap();
#endif
  if (cont) {
    pb.no[p] = next_o; pb.nd[p] = next_d; pb.nh[p] = pack_hit(INF_F, 0xFFFFFFFFu); pb.nslot[p] = slot;
    real_a++;
  }
  }   // tiles
  // null entries for what this warp reserved and did not fill
  for (int k = 0; k < 2; ++k) {
    const uint32_t spare = __shfl_sync(0xffffffffu, out_a.spare, 0) * blk;
    const uint32_t lo = k ? spare : out_a.next, hi = k ? spare + blk : out_a.end;
    if (k && !out_a.has) break;
    for (uint32_t q = lo + lane; q < hi; q += 32u) {
      pb.no[q] = make_float4(0.f, 0.f, 0.f, 0.f); pb.nd[q] = make_float4(0.f, 0.f, 1.f, -1.f);
      pb.nh[q] = NULL_HIT; pb.nslot[q] = 0xFFFFFFFFu;
    }
  }
  for (int k = 0; k < 2; ++k) {
    const uint32_t spare = __shfl_sync(0xffffffffu, out_s.spare, 0) * blk_sh;
    const uint32_t lo = k ? spare : out_s.next, hi = k ? spare + blk_sh : out_s.end;
    if (k && !out_s.has) break;
    for (uint32_t q = lo + lane; q < hi; q += 32u) {
      pb.s_o[q] = make_float4(0.f, 0.f, 0.f, 0.f); pb.s_d[q] = make_float4(0.f, 0.f, 1.f, -1.f);
      pb.s_hits[q] = NULL_HIT; pb.s_contrib[q] = make_float4(0.f, 0.f, 0.f, single ? __uint_as_float(0xFFFFFFFFu) : 0.f);
    }
  }
  real_a = __reduce_add_sync(0xffffffffu, real_a);
  real_s = __reduce_add_sync(0xffffffffu, real_s);
  if (lane == 0) {
    // the lists' lengths in entries (what the traversal and the next bounce read) and the real rays among them
    if (out_a.blocks) atomicAdd(&pb.counts[ACT0 + b + 1], out_a.blocks * blk);
    if (out_s.blocks) atomicAdd(&pb.counts[SH0 + b], out_s.blocks * blk_sh);
    if (real_a) atomicAdd(&pb.counts[REAL0 + b + 1], real_a);
    if (real_s) atomicAdd(&pb.counts[SHV0 + b], real_s);
  }
}

// add the unoccluded light samples in sample order (deterministic), then advance the lists
__global__ void __launch_bounds__(256)
k_resolve_shadow(WaveParams wp, PathBufs pb, uint32_t b) {
  const uint32_t n = pb.counts[ACT0 + b];
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t slot = pb.lslot[i];
  if (slot == 0xFFFFFFFFu) return;   // null entry (k_shade)
  const uint32_t S = wp.S;
  const uint32_t q0 = pb.s_q0[slot];
  if (q0 == 0xFFFFFFFFu) return;
  float4 L = pb.rad[slot];
  bool any = false;
  for (uint32_t j = 0; j < S; ++j) {
    const float4 c = pb.s_contrib[q0 + j];
    if (c.w != 0.f && (uint32_t)pb.s_hits[q0 + j] == 0xFFFFFFFFu) {
      L.x = L.x + c.x; L.y = L.y + c.y; L.z = L.z + c.z;
      any = true;
    }
  }
  if (any) pb.rad[slot] = L;
}

// S == 1 (one light sample per interaction): one thread per SHADOW-list entry; the entry carries its path's slot, a path
// has at most one entry per bounce, so the sum needs no order.  (The path-list walk above reads 8 B per path to find
// the ~55 % of them that own a shadow ray.)
__global__ void __launch_bounds__(256)
k_resolve_shadow1(PathBufs pb, uint32_t b) {
  const uint32_t n = pb.counts[SH0 + b];
  // four entries per thread, all their loads in flight together (one entry per thread left the kernel at half of the
  // HBM bandwidth with 14 % of the issue slots used: too few bytes in flight per SM)
  constexpr int K = 4;
  const uint32_t base = blockIdx.x * (256u * K) + threadIdx.x;
  if (base >= n) return;
  float4 c[K];
  uint32_t prim[K];
#pragma unroll
  for (int k = 0; k < K; ++k) {
    const uint32_t i = base + (uint32_t)k * 256u;
    c[k] = make_float4(0.f, 0.f, 0.f, __uint_as_float(0xFFFFFFFFu)); prim[k] = 0u;
    if (i < n) { c[k] = __ldcs(pb.s_contrib + i); prim[k] = (uint32_t)__ldcs(pb.s_hits + i); }
  }
  float4 L[K];
  bool add[K];
#pragma unroll
  for (int k = 0; k < K; ++k) {
    const uint32_t slot = __float_as_uint(c[k].w);
    add[k] = slot != 0xFFFFFFFFu && prim[k] == 0xFFFFFFFFu;   // not a null entry, not occluded
    if (add[k]) L[k] = pb.rad[slot];
  }
#pragma unroll
  for (int k = 0; k < K; ++k)
    if (add[k]) {
      L[k].x = L[k].x + c[k].x; L[k].y = L[k].y + c[k].y; L[k].z = L[k].z + c[k].z;
      pb.rad[__float_as_uint(c[k].w)] = L[k];
    }
}

// Closes a wave: status 1 = complete, 2 = a ray queue overflowed while it was traced (its paths are incomplete: the
// wave is NOT accumulated and the host re-renders it in halves), 3 = cancelled (b2rt_stop).  Clears the schedulers'
// overflow flags for the next wave.  totals: [0] bounce rays, [1] shadow rays, [2] camera rays of complete waves.
__global__ void k_wave_end(PathBufs pb, uint32_t max_depth, unsigned long long* totals, uint32_t* ctrl_a, uint32_t* ctrl_b,
                           uint32_t n_paths, uint32_t* status) {
  uint32_t ovf = ctrl_a[CTRL_OVERFLOW];
  ctrl_a[CTRL_OVERFLOW] = 0;
  if (ctrl_b) { ovf |= ctrl_b[CTRL_OVERFLOW]; ctrl_b[CTRL_OVERFLOW] = 0; }
  const uint32_t st = pb.counts[3] ? 3u : (ovf ? 2u : 1u);
  *status = st;
  if (st != 1u) return;
  unsigned long long nb = 0, ns = 0;
  for (uint32_t b = 0; b < max_depth; ++b) { if (b) nb += pb.counts[REAL0 + b]; ns += pb.counts[SHV0 + b]; }
  totals[0] += nb; totals[1] += ns; totals[2] += n_paths;
}

// per-pixel accumulation in sample order (deterministic): accum.rgb += L_s, accum.w += 1.  Only complete waves count
// (a cancelled or overflowed wave must not touch the sums or the per-pixel sample count the image is divided by).
__global__ void __launch_bounds__(256)
k_accumulate(WaveParams wp, PathBufs pb, float4* __restrict__ accum, const uint32_t* __restrict__ status) {
  if (*status != 1u) return;
  const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= wp.n_pix) return;
  float4 a = accum[wp.pix0 + p];
  for (uint32_t s = 0; s < wp.spp; ++s) {
    const float4 L = pb.rad[(size_t)p * wp.spp + s];
    a.x = a.x + L.x; a.y = a.y + L.y; a.z = a.z + L.z; a.w = a.w + 1.0f;
  }
  accum[wp.pix0 + p] = a;
}

__global__ void __launch_bounds__(256)
k_resolve_image(const float4* __restrict__ accum, float4* __restrict__ out, uint32_t n_pix, float inv_override) {
  const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n_pix) return;
  const float4 a = accum[p];
  const float inv = inv_override > 0.f ? inv_override : (a.w > 0.f ? __fdiv_rn(1.0f, a.w) : 0.f);
  out[p] = make_float4(a.x * inv, a.y * inv, a.z * inv, 1.0f);
}

// 3x3 per-channel median, out-of-image = 1.0 (kernelMedianFilter, src/cudaRenderer.cu:773-842), on a
// (32+2)x(8+2) shared-memory tile, 19-exchange median-of-9 network.
__device__ __forceinline__ void cswap(float& a, float& b) { const float lo = fminf(a, b), hi = fmaxf(a, b); a = lo; b = hi; }
__device__ __forceinline__ float median9(float v0, float v1, float v2, float v3, float v4, float v5, float v6, float v7, float v8) {
  cswap(v1, v2); cswap(v4, v5); cswap(v7, v8); cswap(v0, v1); cswap(v3, v4); cswap(v6, v7);
  cswap(v1, v2); cswap(v4, v5); cswap(v7, v8); cswap(v0, v3); cswap(v5, v8); cswap(v4, v7);
  cswap(v3, v6); cswap(v1, v4); cswap(v2, v5); cswap(v4, v7); cswap(v4, v2); cswap(v6, v4);
  cswap(v4, v2);
  return v4;
}
__global__ void __launch_bounds__(256)
k_median3x3(const float4* __restrict__ in, float4* __restrict__ out, uint32_t w, uint32_t h) {
  __shared__ float4 tile[10][34];
  const int bx = blockIdx.x * 32, by = blockIdx.y * 8;
  for (int k = threadIdx.y * 32 + threadIdx.x; k < 10 * 34; k += 256) {
    const int ty = k / 34, tx = k % 34;
    const int x = bx + tx - 1, y = by + ty - 1;
    tile[ty][tx] = (x < 0 || y < 0 || x >= (int)w || y >= (int)h) ? make_float4(1.f, 1.f, 1.f, 1.f) : in[(size_t)y * w + x];
  }
  __syncthreads();
  const int x = bx + threadIdx.x, y = by + threadIdx.y;
  if (x >= (int)w || y >= (int)h) return;
  const int tx = threadIdx.x + 1, ty = threadIdx.y + 1;
  float4 r;
#define B2_M(ch) median9(tile[ty - 1][tx - 1].ch, tile[ty - 1][tx].ch, tile[ty - 1][tx + 1].ch, tile[ty][tx - 1].ch, tile[ty][tx].ch, \
                         tile[ty][tx + 1].ch, tile[ty + 1][tx - 1].ch, tile[ty + 1][tx].ch, tile[ty + 1][tx + 1].ch)
  r.x = B2_M(x); r.y = B2_M(y); r.z = B2_M(z); r.w = 1.0f;
#undef B2_M
  out[(size_t)y * w + x] = r;
}

// Gaussian / joint bilateral reconstruction filters (b2rt_config.filter_kind 1 / 2): R = 1 with binomial weights
// (1,2,1) or R = 2 with (1,4,6,4,1); taps outside the image are dropped and the weights renormalised; the bilateral
// range weight is 1 / (1 + |c_q - c_p|^2 * inv_s2).  Taps are accumulated in row-major order with separate multiply and
// add (no contraction), the same order as the oracle's restatement, so the result is bit-identical to it.
template <int R, bool BILATERAL>
__global__ void __launch_bounds__(256)
k_filter(const float4* __restrict__ in, float4* __restrict__ out, uint32_t w, uint32_t h, float inv_s2) {
  constexpr int TW = 32 + 2 * R, TH = 8 + 2 * R;
  __shared__ float4 tile[TH][TW];
  const int bx = blockIdx.x * 32, by = blockIdx.y * 8;
  for (int k = threadIdx.y * 32 + threadIdx.x; k < TH * TW; k += 256) {
    const int ty = k / TW, tx = k % TW;
    const int x = bx + tx - R, y = by + ty - R;
    const bool inside = x >= 0 && y >= 0 && x < (int)w && y < (int)h;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);            // .w = 1 marks a tap inside the image
    if (inside) { v = in[(size_t)y * w + x]; v.w = 1.f; }
    tile[ty][tx] = v;
  }
  __syncthreads();
  const int x = bx + threadIdx.x, y = by + threadIdx.y;
  if (x >= (int)w || y >= (int)h) return;
  const int tx = threadIdx.x + R, ty = threadIdx.y + R;
  const float4 c = tile[ty][tx];
  float ar = 0.f, ag = 0.f, ab = 0.f, ws = 0.f;
#pragma unroll
  for (int dy = -R; dy <= R; ++dy)
#pragma unroll
    for (int dx = -R; dx <= R; ++dx) {
      const float4 q = tile[ty + dy][tx + dx];
      if (q.w == 0.f) continue;
      const float bw = R == 1 ? (float)((2 - (dx < 0 ? -dx : dx)) * (2 - (dy < 0 ? -dy : dy)))
                              : (float)((dx == 0 ? 6 : (dx == 1 || dx == -1 ? 4 : 1)) * (dy == 0 ? 6 : (dy == 1 || dy == -1 ? 4 : 1)));
      float wgt = bw;
      if (BILATERAL) {
        const float dr = q.x - c.x, dg = q.y - c.y, db = q.z - c.z;
        const float d2 = (dr * dr + dg * dg) + db * db;
        wgt = bw * __fdiv_rn(1.0f, 1.0f + d2 * inv_s2);
      }
      ar = ar + wgt * q.x; ag = ag + wgt * q.y; ab = ab + wgt * q.z; ws = ws + wgt;
    }
  out[(size_t)y * w + x] = make_float4(__fdiv_rn(ar, ws), __fdiv_rn(ag, ws), __fdiv_rn(ab, ws), 1.0f);
}

// toColor + update_pixel, src/image.h:49-58,173-188
__global__ void __launch_bounds__(256)
k_tonemap(const float4* __restrict__ img, uint32_t* __restrict__ out, uint32_t n_pix) {
  const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n_pix) return;
  const float4 s = img[p];
  const float one_over_gamma = 1.0f / 2.2f;
  const float exposure = sqrtf(powf(2.f, 1.0f));
  const float r = powf(s.x * exposure, one_over_gamma), g = powf(s.y * exposure, one_over_gamma), b = powf(s.z * exposure, one_over_gamma);
  auto q = [](float c) { c = c < 0.f ? 0.f : (c > 1.f ? 1.f : c); return (uint32_t)(c * 255.f); };
  out[p] = (255u << 24) | (q(b) << 16) | (q(g) << 8) | q(r);
}

__global__ void k_fill_u32(uint32_t* p, uint32_t v) { *p = v; }
// material index -> index | kind << 28: k_shade learns the kind of surface from the word it loads anyway instead of
// from a second, dependent load
__global__ void __launch_bounds__(256)
k_tag_material(uint32_t* __restrict__ prim_material, const b2rt_material* __restrict__ materials, uint32_t n) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t m = prim_material[i];
  prim_material[i] = m | ((uint32_t)materials[m].kind << 28);
}

}  // namespace

// ---- Renderer (host) -------------------------------------------------------------------------------
#define RCHECK(x) do { int rc__ = (x); if (rc__) return rc__; } while (0)

static void free_ptr(void* p) { if (p) cudaFree(p); }

int Renderer::set_device() {
  if (device >= 0) B2RT_CUDA_OK(cudaSetDevice(device));
  return B2RT_OK;
}

int Renderer::create(const b2rt_config* c) {
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    set_error("no CUDA device available (b2rt has no CPU fallback)");
    return B2RT_ERR_NO_DEVICE;
  }
  cfg = *c;
  device = cfg.device;
  if (device < 0) B2RT_CUDA_OK(cudaGetDevice(&device));
  RCHECK(set_device());
  B2RT_CUDA_OK(cudaStreamCreateWithFlags(&own_stream, cudaStreamNonBlocking));
  B2RT_CUDA_OK(cudaStreamCreateWithFlags(&stream_cancel, cudaStreamNonBlocking));
  stream = own_stream;
  B2RT_CUDA_OK(cudaEventCreate(&ev_start));
  B2RT_CUDA_OK(cudaEventCreate(&ev_done));
  B2RT_CUDA_OK(cudaMalloc(&d_sample_base, 4));
  B2RT_CUDA_OK(cudaHostAlloc(&h_sample_base, 4, cudaHostAllocDefault));
  if (const char* e = getenv("B2RT_GRAPH")) graph_off = atoi(e) == 0;
  return B2RT_OK;
}

void Renderer::release_scene() {
  free_bvh(&dbvh);
  free_ptr(d_prim_geom); free_ptr(d_tri_normals_buf); free_ptr(d_prim_material); free_ptr(d_materials); free_ptr(d_lights);
  free_ptr(d_light_area);
  d_prim_geom = nullptr; d_tri_normals = nullptr; d_tri_normals_buf = nullptr; d_prim_material = nullptr; d_materials = nullptr;
  d_lights = nullptr; d_light_area = nullptr;
  cap_prims = cap_normals = cap_mats = cap_lights = 0;
  have_scene = false;
}

void Renderer::release_wave() {
  void* ptrs[] = {l_o[0], l_o[1], l_d[0], l_d[1], l_h[0], l_h[1], l_slot[0], l_slot[1], thr, rad, s_o, s_d, s_hits, s_contrib,
                  s_q0, counts, totals};
  for (void* p : ptrs) free_ptr(p);
  for (int k = 0; k < 2; ++k) { l_o[k] = l_d[k] = nullptr; l_h[k] = nullptr; l_slot[k] = nullptr; }
  thr = rad = s_o = s_d = s_contrib = nullptr; s_hits = nullptr;
  s_q0 = counts = nullptr; totals = nullptr;
  wave_cap = 0;
}

void Renderer::destroy() {
  if (stream) cudaStreamSynchronize(stream);
  if (graph_exec) { cudaGraphExecDestroy(graph_exec); graph_exec = nullptr; }
  free_ptr(d_sample_base); d_sample_base = nullptr;
  if (h_sample_base) { cudaFreeHost(h_sample_base); h_sample_base = nullptr; }
  release_wave();
  tracer.release();
  tracer2.release();
  for (auto& e : ev_sync) cudaEventDestroy(e);
  ev_sync.clear();
  if (stream2) { cudaStreamDestroy(stream2); stream2 = nullptr; }
  release_scene();
  free_ptr(accum); free_ptr(img_a); free_ptr(img_b); free_ptr(ldr); free_ptr(d_env); d_env = nullptr;
  if (host_image) { cudaFreeHost(host_image); host_image = nullptr; host_image_cap = 0; }
  if (ev_start) cudaEventDestroy(ev_start);
  if (ev_done) cudaEventDestroy(ev_done);
  if (own_stream) cudaStreamDestroy(own_stream);
  if (stream_cancel) cudaStreamDestroy(stream_cancel);
  free_ptr(wave_status); wave_status = nullptr; wave_status_cap = 0;
}

int Renderer::set_stream(cudaStream_t s) {
  if (running) RCHECK(wait());
  if (stream) cudaStreamSynchronize(stream);
  stream = s ? s : own_stream;
  return B2RT_OK;
}

int Renderer::set_scene(const b2rt_scene_desc* d) {
  RCHECK(set_device());
  if (running) RCHECK(wait());
  const bool verbose = getenv("B2RT_VERBOSE") != nullptr;
  const auto t0 = std::chrono::steady_clock::now();
  auto lap = [&](const char* what) {
    if (verbose) fprintf(stderr, "b2rt: set_scene %-12s %8.3f ms\n", what, std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count());
  };
  if (!d) { set_error("scene desc is null"); return B2RT_ERR_INVALID; }
  const uint64_t n_prims64 = (uint64_t)d->n_tris + d->n_spheres;
  // cfg.bvh_builder: 0 automatic, 1 host, 2 device; B2RT_BUILDER=gpu|host overrides.  Automatic = the device builder
  // (PLOC) from 2^14 primitives on: its trees trace 1.4 % (cfg3 stand-in) to 3.6 % (cfg2) slower than the host SAH
  // builder's, but 114 K triangles build in 4 ms instead of 22 ms and the host cores are not shared between the ranks
  // of a multi-GPU job (profiles/r02_device_builder_ploc.txt); below that the host build is a few milliseconds.
  bool on_device = cfg.bvh_builder == 2 || (cfg.bvh_builder == 0 && n_prims64 >= (1u << 14));
  if (const char* e = getenv("B2RT_BUILDER")) on_device = !strcmp(e, "gpu");
  on_device = on_device && n_prims64 > 0;
  if (cfg.bvh_builder != 2 && cfg.bvh_width != 0 && cfg.bvh_width != 4 && cfg.bvh_width != 8) on_device = false;   // widths 2 / 16: host builder only
  HostScene hs;
  RCHECK(make_host_scene(d, &hs, !on_device));   // device build: the primitive records are made on the GPU (k_make_prims)
  lap("host scene");
  const size_t np = std::max<size_t>(1, hs.n_prims());
  if (cap_prims < np) {
    free_ptr(d_prim_geom); free_ptr(d_prim_material); d_prim_geom = nullptr; d_prim_material = nullptr;
    cap_prims = np + np / 4;
    B2RT_CUDA_OK(cudaMalloc(&d_prim_geom, cap_prims * PRIM_BYTES));
    B2RT_CUDA_OK(cudaMalloc(&d_prim_material, cap_prims * 4));
  }
  WideBVH wb;
  if (on_device) {
    RCHECK(build_wide_bvh_device(d, cfg.max_leaf_size, cfg.bvh_width, cfg.treelet_bytes, stream, &dbvh, &wb, d_prim_geom));
    lap("bvh build (device)");
  } else {
    RCHECK(build_wide_bvh(hs, cfg.max_leaf_size, cfg.bvh_width, cfg.treelet_bytes, &wb));
    lap("bvh build");
    RCHECK(upload_bvh(wb, &dbvh));   // grow-only device buffers: no cudaMalloc/cudaFree when the scene fits
    lap("bvh upload");
  }
  bvh_stale = true;                // wave buffers are kept; the tracers re-bind their (small) per-subtree arrays
  n_wide_nodes = wb.n_wide_nodes;
  build_ms = wb.build_ms;
  // distance slices (Tracer::trace_sliced): the same automatic rule as b2rt_bvh_build -- on for deep subtree graphs over
  // dense geometry (first slice = 2 mean free paths when that is < 1/8 of the scene diagonal), off for the box scenes
  for (int k = 0; k < 6; ++k) tracer.slice_bbox[k] = wb.bbox[k];
  tracer.slice_first = 0.f; tracer.slice_growth = 4.f; tracer.slice_passes = 4;
  {
    const float ex = wb.bbox[3] - wb.bbox[0], ey = wb.bbox[4] - wb.bbox[1], ez = wb.bbox[5] - wb.bbox[2];
    const float want = 2.f * wb.mean_free_path;
    if (wb.n_levels >= 3 && want > 0.f && want < 0.125f * std::sqrt(ex * ex + ey * ey + ez * ez)) tracer.slice_first = want;
  }
  if (const char* e = getenv("B2RT_RENDER_SLICE")) {   // experiment: distance-sliced traversal inside the renderer
    float f = 0.f, g = 4.f; int p = 2;
    if (sscanf(e, "%f,%f,%d", &f, &g, &p) >= 1) { tracer.slice_first = f; tracer.slice_growth = g > 1.f ? g : 4.f; tracer.slice_passes = p >= 2 ? p : 2; }
  }
  n_tris = hs.n_tris;
  n_lights = (uint32_t)hs.lights.size();
  if (cap_mats < hs.materials.size()) {
    free_ptr(d_materials); d_materials = nullptr;
    cap_mats = hs.materials.size() + 16;
    B2RT_CUDA_OK(cudaMalloc(&d_materials, cap_mats * sizeof(b2rt_material)));
  }
  if (cap_lights < std::max<size_t>(1, hs.lights.size())) {
    free_ptr(d_lights); free_ptr(d_light_area); d_lights = nullptr; d_light_area = nullptr;
    cap_lights = hs.lights.size() + 8;
    B2RT_CUDA_OK(cudaMalloc(&d_lights, cap_lights * sizeof(b2rt_light)));
    B2RT_CUDA_OK(cudaMalloc(&d_light_area, cap_lights * 4));
  }
  if (hs.n_prims()) {
    if (!on_device) B2RT_CUDA_OK(cudaMemcpy(d_prim_geom, hs.prim_geom.data(), (size_t)hs.n_prims() * PRIM_BYTES, cudaMemcpyHostToDevice));
    B2RT_CUDA_OK(cudaMemcpy(d_prim_material, hs.prim_material.data(), (size_t)hs.n_prims() * 4, cudaMemcpyHostToDevice));
  }
  B2RT_CUDA_OK(cudaMemcpy(d_materials, hs.materials.data(), hs.materials.size() * sizeof(b2rt_material), cudaMemcpyHostToDevice));
  if (hs.n_prims()) {
    k_tag_material<<<(uint32_t)((hs.n_prims() + 255) / 256), 256, 0, stream>>>(d_prim_material, d_materials, (uint32_t)hs.n_prims());
    B2RT_CUDA_OK(cudaGetLastError());
  }
  if (d->tri_normals && d->n_tris) {   // straight from the caller's array
    const size_t nn = (size_t)d->n_tris * 9;
    if (cap_normals < nn) {
      free_ptr(d_tri_normals_buf); d_tri_normals_buf = nullptr;
      cap_normals = nn + nn / 4;
      B2RT_CUDA_OK(cudaMalloc(&d_tri_normals_buf, cap_normals * 4));
    }
    d_tri_normals = d_tri_normals_buf;
    B2RT_CUDA_OK(cudaMemcpy(d_tri_normals, d->tri_normals, nn * 4, cudaMemcpyHostToDevice));
  } else {
    d_tri_normals = nullptr;
  }
  if (!hs.lights.empty()) {
    std::vector<float> area(hs.lights.size());
    for (size_t i = 0; i < hs.lights.size(); ++i) {
      const b2rt_light& l = hs.lights[i];
      // fp32, same expression as the oracle: |dim_x| * |dim_y| with dot = fma(z,z,fma(y,y,x*x))
      auto len = [](const float* v) { return sqrtf(fmaf(v[2], v[2], fmaf(v[1], v[1], v[0] * v[0]))); };
      area[i] = len(l.dim_x) * len(l.dim_y);
    }
    B2RT_CUDA_OK(cudaMemcpy(d_lights, hs.lights.data(), hs.lights.size() * sizeof(b2rt_light), cudaMemcpyHostToDevice));
    B2RT_CUDA_OK(cudaMemcpy(d_light_area, area.data(), area.size() * 4, cudaMemcpyHostToDevice));
  }
  // shadow rays per interaction
  shadow_per_hit = 0;
  for (auto& l : hs.lights) shadow_per_hit += l.kind == B2RT_LIGHT_AREA ? std::max(1u, cfg.ns_area_light) : 1u;
  lights_host = hs.lights;
  have_glossy = false;
  for (auto& m : hs.materials) have_glossy = have_glossy || m.kind == B2RT_MAT_GLOSSY;
  have_scene = true;
  B2RT_CUDA_OK(cudaDeviceSynchronize());   // uploads above used the legacy stream; work runs on `stream`
  lap("scene upload");
  return B2RT_OK;
}

int Renderer::set_envmap(const float* rgb, uint32_t w, uint32_t h) {
  RCHECK(set_device());
  if (running) RCHECK(wait());
  B2RT_CUDA_OK(cudaStreamSynchronize(stream));
  free_ptr(d_env); d_env = nullptr; env_w = env_h = 0;
  if (rgb) {
    if (!w || !h || (uint64_t)w * h > (1ull << 28)) { set_error("invalid environment map size"); return B2RT_ERR_INVALID; }
    B2RT_CUDA_OK(cudaMalloc(&d_env, (size_t)w * h * 12));
    B2RT_CUDA_OK(cudaMemcpy(d_env, rgb, (size_t)w * h * 12, cudaMemcpyHostToDevice));
    env_w = w; env_h = h;
  }
  return clear();
}

int Renderer::set_camera(const b2rt_camera* c) {
  cam = *c;
  have_camera = true;
  return clear();   // a new viewpoint restarts accumulation (CudaRenderer::setViewpoint, cudaRenderer.cu:1866-1869)
}

int Renderer::set_frame_size(uint32_t w, uint32_t h) {
  if (w == 0 || h == 0 || (uint64_t)w * h > 0x7FFFFFFFull) { set_error("invalid frame size"); return B2RT_ERR_INVALID; }
  RCHECK(set_device());
  if (running) RCHECK(wait());
  if (w == width && h == height && accum) return clear();
  free_ptr(accum); free_ptr(img_a); free_ptr(img_b); free_ptr(ldr);
  accum = img_a = img_b = nullptr; ldr = nullptr;
  width = w; height = h;
  const size_t np = (size_t)w * h;
  B2RT_CUDA_OK(cudaMalloc(&accum, np * sizeof(float4)));
  B2RT_CUDA_OK(cudaMalloc(&img_a, np * sizeof(float4)));
  B2RT_CUDA_OK(cudaMalloc(&img_b, np * sizeof(float4)));
  B2RT_CUDA_OK(cudaMalloc(&ldr, np * 4));
  release_wave();
  return clear();
}

int Renderer::clear() {
  RCHECK(set_device());
  if (running) RCHECK(wait());
  if (accum) B2RT_CUDA_OK(cudaMemsetAsync(accum, 0, (size_t)width * height * sizeof(float4), stream));
  samples_done = 0;
  return B2RT_OK;
}

int Renderer::ensure_wave() {
  // shadow rays per interaction can change with the knobs
  uint32_t S = 0;
  for (auto& l : lights_host) S += l.kind == B2RT_LIGHT_AREA ? std::max(1u, cfg.ns_area_light) : 1u;
  if (d_env) S += 1;   // the environment map is sampled like one more light
  shadow_per_hit = S;
  const uint64_t n_pix = (uint64_t)width * height;
  uint64_t cap = cfg.max_wave_paths ? cfg.max_wave_paths : (64u << 20);   // ~180 B of state per path + two schedulers: ~19 GB of the 180 GB
  cap = std::max<uint64_t>(cap, 1024);
  const uint64_t want = std::min<uint64_t>(cap, n_pix * std::max(1u, cfg.ns_aa));
  const uint32_t Salloc = std::max(1u, S);
  if (!shade_ctas) {
    int sms = 0;
    B2RT_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
    shade_ctas = (uint32_t)sms * B2RT_SHADE_OCC;
  }
  // k_shade's null entries: every warp may leave up to two reservation blocks unfilled per list and bounce
  const uint64_t slack = (uint64_t)shade_ctas * (SHADE_THREADS / 32) * 2 * SHADE_BLK_MAX;
  list_slack = (uint32_t)slack;
  // Two schedulers on two streams (default; B2RT_OVERLAP=0 turns it off): see start().
  overlap = getenv("B2RT_OVERLAP") ? atoi(getenv("B2RT_OVERLAP")) != 0 : true;
  if (overlap && !stream2) {
    B2RT_CUDA_OK(cudaStreamCreateWithFlags(&stream2, cudaStreamNonBlocking));
    ev_sync.resize(2 * MAX_DEPTH);
    for (auto& e : ev_sync) B2RT_CUDA_OK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  }
  if (wave_cap >= want && wave_S >= Salloc && tracer.max_rays >= (want + slack) * Salloc && (!overlap || tracer2.max_rays >= (want + slack) * Salloc)) {
    if (bvh_stale) {
      RCHECK(tracer.init(dbvh, tracer.max_rays, pair_factor));
      if (overlap) RCHECK(tracer2.init(dbvh, tracer2.max_rays, pair_factor));
      bvh_stale = false;
    }
    return B2RT_OK;
  }
  release_wave();
  tracer.release();
  tracer2.release();
  wave_cap = want; wave_S = Salloc;
  // (+16 entries: the traversal copies ray tiles in 16-byte units and may read up to 3 entries past a list's end)
  for (int k = 0; k < 2; ++k) {
    B2RT_CUDA_OK(cudaMalloc(&l_o[k], (wave_cap + slack + 16) * sizeof(float4)));
    B2RT_CUDA_OK(cudaMalloc(&l_d[k], (wave_cap + slack + 16) * sizeof(float4)));
    B2RT_CUDA_OK(cudaMalloc(&l_h[k], (wave_cap + slack + 16) * 8));
    B2RT_CUDA_OK(cudaMalloc(&l_slot[k], (wave_cap + slack + 16) * 4));
  }
  B2RT_CUDA_OK(cudaMalloc(&thr, wave_cap * sizeof(float4)));
  B2RT_CUDA_OK(cudaMalloc(&rad, wave_cap * sizeof(float4)));
  B2RT_CUDA_OK(cudaMalloc(&s_o, ((wave_cap + slack) * Salloc + 16) * sizeof(float4)));
  B2RT_CUDA_OK(cudaMalloc(&s_d, ((wave_cap + slack) * Salloc + 16) * sizeof(float4)));
  B2RT_CUDA_OK(cudaMalloc(&s_hits, ((wave_cap + slack) * Salloc + 16) * 8));
  B2RT_CUDA_OK(cudaMalloc(&s_contrib, (wave_cap + slack) * Salloc * sizeof(float4)));
  B2RT_CUDA_OK(cudaMalloc(&s_q0, wave_cap * 4));
  B2RT_CUDA_OK(cudaMalloc(&counts, N_COUNTS * 4));
  B2RT_CUDA_OK(cudaMalloc(&totals, 8 * 8));
  B2RT_CUDA_OK(cudaMemset(counts, 0, N_COUNTS * 4));
  B2RT_CUDA_OK(cudaMemset(totals, 0, 8 * 8));
  if (const char* e = getenv("B2RT_PAIR_FACTOR")) {   // starting value (tests: 1 forces the growth path on a push-heavy scene)
    int v = atoi(e);
    if (v >= 1 && v <= 64 && !pair_factor_from_env) { pair_factor = (uint32_t)v; pair_factor_from_env = true; }
  }
  RCHECK(tracer.init(dbvh, (wave_cap + slack) * Salloc, pair_factor));
  if (overlap) RCHECK(tracer2.init(dbvh, (wave_cap + slack) * Salloc, pair_factor));
  bvh_stale = false;
  return B2RT_OK;
}

// host copy of what a wave needs besides its WaveParams (same for every wave of a frame)
struct Renderer::FrameCtx {
  CamDev cd; SceneDev sd; PathBufs pb; uint32_t max_depth; uint32_t S;
};

int Renderer::make_frame_ctx(FrameCtx* fc) {
  fc->max_depth = std::min(MAX_DEPTH, std::max(1u, cfg.max_ray_depth));
  fc->S = shadow_per_hit;
  CamDev& cd = fc->cd;
  cd.pos = f3{cam.pos[0], cam.pos[1], cam.pos[2]};
  cd.cx = f3{cam.c2w[0], cam.c2w[1], cam.c2w[2]};
  cd.cy = f3{cam.c2w[3], cam.c2w[4], cam.c2w[5]};
  cd.cz = f3{cam.c2w[6], cam.c2w[7], cam.c2w[8]};
  cd.tan_h = tanf(cam.hfov_deg * 0.5f * 0.01745329251994329577f);
  cd.tan_v = tanf(cam.vfov_deg * 0.5f * 0.01745329251994329577f);
  SceneDev& sd = fc->sd;
  sd.prim_geom = (const float4*)d_prim_geom; sd.tri_normals = d_tri_normals; sd.prim_material = d_prim_material;
  sd.materials = d_materials; sd.lights = d_lights; sd.light_area = d_light_area; sd.n_tris = n_tris; sd.n_lights = n_lights;
  sd.env = d_env; sd.env_w = d_env ? env_w : 0u; sd.env_h = d_env ? env_h : 0u;
  PathBufs& pb = fc->pb;
  pb.lo = (float4*)l_o[0]; pb.ld = (float4*)l_d[0]; pb.lh = l_h[0]; pb.lslot = l_slot[0];
  pb.no = (float4*)l_o[1]; pb.nd = (float4*)l_d[1]; pb.nh = l_h[1]; pb.nslot = l_slot[1];
  pb.thr = (float4*)thr; pb.rad = (float4*)rad;
  pb.s_o = (float4*)s_o; pb.s_d = (float4*)s_d; pb.s_hits = s_hits; pb.s_contrib = (float4*)s_contrib;
  pb.s_q0 = s_q0; pb.counts = counts; pb.inc_bound = 0xFFFFFFFFu;
  pb.list_cap = (uint32_t)std::min<uint64_t>(wave_cap + list_slack, 0xFFFFFFFFull);
  pb.shadow_cap = (uint32_t)std::min<uint64_t>((wave_cap + list_slack) * std::max(1u, wave_S), 0xFFFFFFFFull);
  return B2RT_OK;
}

// Enqueues one wave (ray generation, max_depth bounces, wave end, accumulation) on the renderer's streams; the wave's
// outcome lands in wave_status[status_index] (k_wave_end).
int Renderer::enqueue_wave(const FrameCtx& fc, const WaveParams& wp, uint32_t status_index) {
  PathBufs pb = fc.pb;
  auto bind_lists = [&](uint32_t cur) {   // list `cur` is read, the other one is appended to
    const uint32_t nxt = cur ^ 1u;
    pb.lo = (float4*)l_o[cur]; pb.ld = (float4*)l_d[cur]; pb.lh = l_h[cur]; pb.lslot = l_slot[cur];
    pb.no = (float4*)l_o[nxt]; pb.nd = (float4*)l_d[nxt]; pb.nh = l_h[nxt]; pb.nslot = l_slot[nxt];
  };
  const uint32_t max_depth = fc.max_depth, S = fc.S;
  const bool shade_ext = d_env != nullptr || have_glossy;
  const uint32_t n = wp.n_pix * wp.spp;
  const uint32_t g = (n + 255) / 256;
  const uint64_t n_list = (uint64_t)n + list_slack;       // longest list of the wave, null entries included
  const uint32_t g_list = (uint32_t)((n_list + 255) / 256);
  bind_lists(0);   // k_raygen fills list 0; k_shade(b) appends the continuing paths to list (b+1)&1
  k_raygen<<<g, 256, 0, stream>>>(wp, fc.cd, pb); launches++;
  // The shadow rays of bounce b and the continuing rays of bounce b + 1 both come out of k_shade(b) and do not
  // depend on each other: the any-hit trace + k_resolve_shadow(b) run on a second stream with a second scheduler
  // next to the closest-hit trace of bounce b + 1, which fills the issue slots the level >= 1 launches and every
  // launch's tail leave idle (cfg2 -4.9 %, cfg3 stand-in -5.2 % per frame, tools/ab_overlap.sh).  Not while
  // per-launch timing is on (b2rt_set_profiling): launches are then timed alone, on one stream.
  const bool ov = overlap && !tracer.time_kernels;
  bool pending_resolve = false;
  for (uint32_t b = 0; b < max_depth; ++b) {
    bind_lists(b & 1u);
    RCHECK(tracer.trace_sliced(stream, pb.lo, pb.ld, pb.lh, counts + ACT0 + b, n_list, false));
    // shade(b) adds emission to the radiance that resolve(b - 1) updates and rewrites the shadow list it reads
    if (pending_resolve) { B2RT_CUDA_OK(cudaStreamWaitEvent(stream, ev_sync[2 * (b - 1) + 1], 0)); pending_resolve = false; }
    if (shade_ext) k_shade<true><<<shade_ctas, SHADE_THREADS, 0, stream>>>(wp, fc.sd, pb, b);
    else k_shade<false><<<shade_ctas, SHADE_THREADS, 0, stream>>>(wp, fc.sd, pb, b);
    launches++;
    if (S > 0) {
      if (ov) {
        // shadow rays of bounce b on the second stream, next to the closest-hit trace of bounce b + 1
        B2RT_CUDA_OK(cudaEventRecord(ev_sync[2 * b], stream));
        B2RT_CUDA_OK(cudaStreamWaitEvent(stream2, ev_sync[2 * b], 0));
        RCHECK(tracer2.trace_sliced(stream2, pb.s_o, pb.s_d, pb.s_hits, counts + SH0 + b, n_list * S, true));
        if (S == 1) k_resolve_shadow1<<<(g_list + 3) / 4, 256, 0, stream2>>>(pb, b);
        else k_resolve_shadow<<<g_list, 256, 0, stream2>>>(wp, pb, b);
        launches++;
        B2RT_CUDA_OK(cudaEventRecord(ev_sync[2 * b + 1], stream2));
        pending_resolve = true;
      } else {
        RCHECK(tracer.trace_sliced(stream, pb.s_o, pb.s_d, pb.s_hits, counts + SH0 + b, n_list * S, true));
        if (S == 1) k_resolve_shadow1<<<(g_list + 3) / 4, 256, 0, stream>>>(pb, b);
        else k_resolve_shadow<<<g_list, 256, 0, stream>>>(wp, pb, b);
        launches++;
      }
    }
  }
  if (pending_resolve) B2RT_CUDA_OK(cudaStreamWaitEvent(stream, ev_sync[2 * (max_depth - 1) + 1], 0));
  k_wave_end<<<1, 1, 0, stream>>>(pb, max_depth, totals, tracer.ctrl, (ov && tracer2.ctrl) ? tracer2.ctrl : nullptr, n,
                                  wave_status + status_index); launches++;
  k_accumulate<<<(wp.n_pix + 255) / 256, 256, 0, stream>>>(wp, pb, (float4*)accum, wave_status + status_index); launches++;
  return B2RT_OK;
}

int Renderer::start() {
  if (!have_scene || !have_camera || !accum) { set_error("start: scene, camera and frame size must be set first"); return B2RT_ERR_INVALID; }
  if (cfg.ns_aa == 0) { set_error("ns_aa must be >= 1"); return B2RT_ERR_INVALID; }
  RCHECK(set_device());
  if (running) RCHECK(wait());
  RCHECK(ensure_wave());
  const uint32_t S = shadow_per_hit;
  const uint32_t stride = cfg.sample_stride ? cfg.sample_stride : 1;
  const uint64_t n_pix = (uint64_t)width * height;
  FrameCtx fc;
  RCHECK(make_frame_ctx(&fc));

  tracer.launches = 0; tracer.traverse_launches = 0; tracer.traverse_launches_l0 = 0; tracer.ev_used = 0; tracer.ev_deeper.clear();
  tracer2.launches = 0; tracer2.traverse_launches = 0; tracer2.traverse_launches_l0 = 0; tracer2.ev_used = 0; tracer2.ev_deeper.clear();
  tracer2.collect_stats = tracer.collect_stats; tracer2.time_kernels = tracer.time_kernels;
  tracer2.slice_first = tracer.slice_first; tracer2.slice_growth = tracer.slice_growth; tracer2.slice_passes = tracer.slice_passes;
  for (int k = 0; k < 6; ++k) tracer2.slice_bbox[k] = tracer.slice_bbox[k];
  if (overlap) RCHECK(tracer2.reset_counters(stream));
  launches = 0;
  B2RT_CUDA_OK(cudaMemsetAsync(totals, 0, 8 * 8, stream));
  RCHECK(tracer.reset_counters(stream));
  B2RT_CUDA_OK(cudaMemsetAsync(counts + 3, 0, 4, stream));  // cancel flag
  B2RT_CUDA_OK(cudaEventRecord(ev_start, stream));
  ms_traverse_acc = 0;

  // waves: pixel ranges x sample chunks, ascending in samples so accumulation order is fixed
  uint32_t spp_chunk, pix_chunk;
  if (n_pix >= wave_cap) { spp_chunk = 1; pix_chunk = (uint32_t)wave_cap; }
  else { spp_chunk = (uint32_t)std::min<uint64_t>(cfg.ns_aa, wave_cap / n_pix); pix_chunk = (uint32_t)n_pix; }
  waves.clear();
  for (uint32_t s0 = 0; s0 < cfg.ns_aa; s0 += spp_chunk) {
    const uint32_t spp = std::min(spp_chunk, cfg.ns_aa - s0);
    for (uint64_t p0 = 0; p0 < n_pix; p0 += pix_chunk) {
      WaveParams wp;
      wp.pix0 = (uint32_t)p0; wp.n_pix = (uint32_t)std::min<uint64_t>(pix_chunk, n_pix - p0);
      wp.spp = spp; wp.sample0 = s0 * stride; wp.sample_base = d_sample_base; wp.sample_stride = stride;
      wp.width = width; wp.height = height;
      wp.jitter = (cfg.ns_aa * stride) > 1 ? 1u : 0u;
      wp.k0 = (uint32_t)cfg.seed; wp.k1 = (uint32_t)(cfg.seed >> 32);
      wp.eps = cfg.ray_eps > 0.f ? cfg.ray_eps : 1e-4f;
      wp.max_depth = fc.max_depth; wp.ns_area_light = cfg.ns_area_light; wp.S = S;
      waves.push_back(wp);
    }
  }
  // one status word per wave (+ one scratch word for re-rendered halves)
  if (wave_status_cap < waves.size() + 1) {
    free_ptr(wave_status); wave_status = nullptr;
    wave_status_cap = waves.size() + 1 + 64;
    B2RT_CUDA_OK(cudaMalloc(&wave_status, wave_status_cap * 4));
  }
  B2RT_CUDA_OK(cudaMemsetAsync(wave_status, 0, (waves.size() + 1) * 4, stream));
  *h_sample_base = cfg.sample_first;
  B2RT_CUDA_OK(cudaMemcpyAsync(d_sample_base, h_sample_base, 4, cudaMemcpyHostToDevice, stream));
  // eager, capture or replay (see render.cuh)
  const bool may_graph = !graph_off && !tracer.time_kernels;
  const uint64_t sig = may_graph ? frame_signature(fc) : 0;
  if (may_graph && graph_exec && sig == graph_sig) {
    B2RT_CUDA_OK(cudaGraphLaunch(graph_exec, stream));
    launches = g_launches;
    tracer.launches = g_t1[0]; tracer.traverse_launches = g_t1[1]; tracer.traverse_launches_l0 = g_t1[2];
    tracer2.launches = g_t2[0]; tracer2.traverse_launches = g_t2[1]; tracer2.traverse_launches_l0 = g_t2[2];
    graph_replays++;
  } else if (may_graph && sig == last_sig) {
    // second identical frame in a row: capture it (every buffer it needs exists since the first one), then launch it
    if (graph_exec) { cudaGraphExecDestroy(graph_exec); graph_exec = nullptr; }
    cudaGraph_t g = nullptr;
    int rc = B2RT_OK;
    if (cudaStreamBeginCapture(stream, cudaStreamCaptureModeThreadLocal) == cudaSuccess) {
      for (size_t w = 0; w < waves.size() && rc == B2RT_OK; ++w) rc = enqueue_wave(fc, waves[w], (uint32_t)w);
      const cudaError_t ce = cudaStreamEndCapture(stream, &g);
      if (rc == B2RT_OK && ce == cudaSuccess && g && cudaGraphInstantiate(&graph_exec, g, 0) == cudaSuccess) {
        graph_sig = sig;
        g_launches = launches;
        g_t1[0] = tracer.launches; g_t1[1] = tracer.traverse_launches; g_t1[2] = tracer.traverse_launches_l0;
        g_t2[0] = tracer2.launches; g_t2[1] = tracer2.traverse_launches; g_t2[2] = tracer2.traverse_launches_l0;
      } else {
        graph_exec = nullptr;
      }
      if (g) cudaGraphDestroy(g);
    }
    cudaGetLastError();   // (a failed capture leaves a sticky-looking error code behind; the eager path below is the answer to it)
    if (graph_exec) {
      B2RT_CUDA_OK(cudaGraphLaunch(graph_exec, stream));
    } else {
      graph_off = true;   // this renderer's frames cannot be captured: stay eager
      launches = 0;
      tracer.launches = tracer.traverse_launches = tracer.traverse_launches_l0 = 0;
      tracer2.launches = tracer2.traverse_launches = tracer2.traverse_launches_l0 = 0;
      for (size_t w = 0; w < waves.size(); ++w) RCHECK(enqueue_wave(fc, waves[w], (uint32_t)w));
    }
  } else {
    for (size_t w = 0; w < waves.size(); ++w) RCHECK(enqueue_wave(fc, waves[w], (uint32_t)w));
  }
  last_sig = sig;
  B2RT_CUDA_OK(cudaGetLastError());
  B2RT_CUDA_OK(cudaEventRecord(ev_done, stream));
  running = true;
  return B2RT_OK;
}

// Everything a frame's launch sequence depends on besides the contents of device memory: the frame context (camera, scene
// and list pointers, counts), the waves, both schedulers' buffers and launch shapes, the subtree levels, the mode flags.
uint64_t Renderer::frame_signature(const FrameCtx& fc) const {
  uint64_t h = 1469598103934665603ull;
  auto mix = [&](const void* p, size_t n) { const uint8_t* b = (const uint8_t*)p; for (size_t i = 0; i < n; ++i) { h ^= b[i]; h *= 1099511628211ull; } };
  auto mixv = [&](uint64_t v) { mix(&v, 8); };
  // (struct padding is not hashed: field by field)
  mix(&fc.cd.pos, sizeof fc.cd.pos); mix(&fc.cd.cx, sizeof fc.cd.cx); mix(&fc.cd.cy, sizeof fc.cd.cy); mix(&fc.cd.cz, sizeof fc.cd.cz);
  mix(&fc.cd.tan_h, 4); mix(&fc.cd.tan_v, 4);
  const SceneDev& sd = fc.sd;
  mixv((uint64_t)sd.prim_geom); mixv((uint64_t)sd.tri_normals); mixv((uint64_t)sd.prim_material); mixv((uint64_t)sd.materials);
  mixv((uint64_t)sd.lights); mixv((uint64_t)sd.light_area); mixv(sd.n_tris); mixv(sd.n_lights); mixv((uint64_t)sd.env); mixv(sd.env_w); mixv(sd.env_h);
  const PathBufs& pb = fc.pb;
  const void* ptrs[] = {pb.lo, pb.ld, pb.lh, pb.lslot, pb.no, pb.nd, pb.nh, pb.nslot, pb.thr, pb.rad, pb.s_o, pb.s_d, pb.s_hits, pb.s_contrib,
                        pb.s_q0, pb.counts, accum, wave_status, totals, d_sample_base, (const void*)stream, (const void*)stream2};
  for (const void* q : ptrs) mixv((uint64_t)q);
  mixv(pb.list_cap); mixv(pb.shadow_cap); mixv(fc.max_depth); mixv(fc.S); mixv(overlap); mixv(have_glossy); mixv(shade_ctas); mixv(list_slack);
  for (const WaveParams& w : waves) {
    const uint32_t v[] = {w.pix0, w.n_pix, w.spp, w.sample0, w.sample_stride, w.width, w.height, w.jitter, w.k0, w.k1, w.max_depth, w.ns_area_light, w.S};
    mix(v, sizeof v); mix(&w.eps, 4);
  }
  for (const Tracer* t : {&tracer, &tracer2}) {
    const void* tp[] = {t->bvh.blob, t->bvh.treelets, t->cnt, t->seg_off, t->cursor, t->pairs, t->ids_sorted, t->chunks, t->ctrl, t->counters,
                        t->sched_scratch, t->sl_n, t->sl_o[0], t->sl_o[1]};
    for (const void* q : tp) mixv((uint64_t)q);
    mixv(t->pair_cap); mixv(t->max_rays); mixv(t->chunk_cap); mixv(t->chunk_rays); mixv(t->chunk_min); mixv(t->chunks_per_cta); mixv(t->chunk0_max);
    mixv(t->smem_bytes); mixv(t->stack_off); mixv(t->ctas_per_sm); mixv(t->num_sms); mixv(t->count_ctas); mixv(t->scatter_ctas);
    mixv(t->collect_stats); mixv(t->time_kernels); mixv(t->slice_passes); mix(&t->slice_first, 4); mix(&t->slice_growth, 4); mix(t->slice_bbox, sizeof t->slice_bbox);
    mixv(t->bvh.n_treelets); mixv(t->bvh.n_levels); mixv(t->bvh.width); mixv(t->bvh.max_treelet_bytes);
    for (uint32_t L = 0; L < t->bvh.n_levels; ++L) { mixv(t->bvh.levels[L].first); mixv(t->bvh.levels[L].count); }
  }
  return h ? h : 1;
}

int Renderer::is_done() {
  if (!running) return 1;
  cudaError_t e = cudaEventQuery(ev_done);
  if (e == cudaSuccess) { int rc = wait(); return rc ? rc : 1; }
  if (e == cudaErrorNotReady) return 0;
  set_error(std::string("cudaEventQuery: ") + cudaGetErrorString(e));
  return B2RT_ERR_CUDA;
}

// A wave whose ray queues overflowed was not accumulated: render it again as two halves (samples, or pixels when it
// has one sample), recursively, blocking.  The halves reuse the frame's buffers, whose queues were sized for the
// whole wave.  Its samples are then added after those of the later waves, so the fp32 sums of such a frame can differ
// in the last bits from a frame that never overflowed.
int Renderer::retry_wave(const FrameCtx& fc, const WaveParams& wp, int depth) {
  const uint64_t n = (uint64_t)wp.n_pix * wp.spp;
  if (n < 2048 || depth > 24) { set_error("ray queue overflow on a minimal wave"); return B2RT_ERR_OVERFLOW; }
  WaveParams half[2] = {wp, wp};
  if (wp.spp > 1) {
    half[0].spp = wp.spp / 2; half[1].spp = wp.spp - half[0].spp;
    half[1].sample0 = wp.sample0 + half[0].spp * wp.sample_stride;
  } else {
    half[0].n_pix = wp.n_pix / 2; half[1].n_pix = wp.n_pix - half[0].n_pix;
    half[1].pix0 = wp.pix0 + half[0].n_pix;
  }
  const uint32_t scratch = (uint32_t)waves.size();
  for (int k = 0; k < 2; ++k) {
    RCHECK(enqueue_wave(fc, half[k], scratch));
    uint32_t st = 0;
    B2RT_CUDA_OK(cudaMemcpyAsync(&st, wave_status + scratch, 4, cudaMemcpyDeviceToHost, stream));
    B2RT_CUDA_OK(cudaStreamSynchronize(stream));
    if (st == 2u) RCHECK(retry_wave(fc, half[k], depth + 1));
    else if (st == 1u) waves_retried++;
  }
  return B2RT_OK;
}

int Renderer::wait() {
  if (!running) return B2RT_OK;
  RCHECK(set_device());
  B2RT_CUDA_OK(cudaEventSynchronize(ev_done));
  running = false;
  float ms = 0;
  B2RT_CUDA_OK(cudaEventElapsedTime(&ms, ev_start, ev_done));
  ms_total = ms;
  // outcome of every wave; overflowed waves are rendered again in halves before anything is reported
  std::vector<uint32_t> st(waves.size());
  if (!st.empty()) B2RT_CUDA_OK(cudaMemcpy(st.data(), wave_status, st.size() * 4, cudaMemcpyDeviceToHost));
  bool cancelled = false;
  waves_retried = 0; queues_grown = 0;
  FrameCtx fc;
  bool have_fc = false;
  for (size_t w = 0; w < st.size(); ++w) {
    if (st[w] == 3u) cancelled = true;
    if (st[w] != 2u) continue;
    if (!have_fc) { RCHECK(make_frame_ctx(&fc)); have_fc = true; }
    // The queues were sized for pair_factor pushes per ray and level (4 covers the box scenes' 0.3-0.5 and the soup's
    // 2-3).  A scene that pushes more gets larger queues -- for this wave and every later frame -- before anything is
    // split: the stream is idle here, so the schedulers can be re-initialised in place.
    uint32_t st_w = 2u;
    while (st_w == 2u && pair_factor < 32 && !getenv("B2RT_DEBUG_PAIR_CAP")) {
      pair_factor *= 2;
      if (tracer.init(dbvh, tracer.max_rays, pair_factor) != B2RT_OK || (overlap && tracer2.init(dbvh, tracer2.max_rays, pair_factor) != B2RT_OK)) {
        // no memory for larger queues: go back and split the wave instead
        pair_factor /= 2;
        RCHECK(tracer.init(dbvh, tracer.max_rays, pair_factor));
        if (overlap) RCHECK(tracer2.init(dbvh, tracer2.max_rays, pair_factor));
        break;
      }
      queues_grown++;
      const uint32_t scratch = (uint32_t)waves.size();
      RCHECK(enqueue_wave(fc, waves[w], scratch));
      B2RT_CUDA_OK(cudaMemcpyAsync(&st_w, wave_status + scratch, 4, cudaMemcpyDeviceToHost, stream));
      B2RT_CUDA_OK(cudaStreamSynchronize(stream));
      if (st_w == 1u) waves_retried++;
    }
    if (st_w == 2u) RCHECK(retry_wave(fc, waves[w], 0));
  }
  unsigned long long t[8];
  B2RT_CUDA_OK(cudaMemcpy(t, totals, sizeof t, cudaMemcpyDeviceToHost));
  // samples per pixel this call added: the frame's ns_aa, or (after b2rt_stop) the average over the pixels
  const uint64_t n_pix = (uint64_t)width * height;
  samples_done += cancelled ? (n_pix ? t[2] / n_pix : 0) : cfg.ns_aa;
  TraceCounters tc, tc0;
  RCHECK(tracer.read_counters(&tc, &tc0));
  last = b2rt_stats();
  last.rays_camera = t[2];
  last.rays_bounce = t[0]; last.rays_shadow = t[1];
  last.node_visits = tc.node_visits; last.leaf_prim_tests = tc.prim_tests; last.subtree_visits = tc.subtree_visits;
  last.queue_pushes = tc.pushes; last.staged_bytes = tc.staged_bytes; last.hit_updates = tc.hit_updates;
  last.node_visits_l0 = tc0.node_visits; last.leaf_prim_tests_l0 = tc0.prim_tests; last.queue_pushes_l0 = tc0.pushes;
  last.staged_bytes_l0 = tc0.staged_bytes; last.hit_updates_l0 = tc0.hit_updates;
  if (overlap && tracer2.counters) {
    TraceCounters t2, t20;
    RCHECK(tracer2.read_counters(&t2, &t20));
    last.node_visits += t2.node_visits; last.leaf_prim_tests += t2.prim_tests; last.subtree_visits += t2.subtree_visits;
    last.queue_pushes += t2.pushes; last.staged_bytes += t2.staged_bytes; last.hit_updates += t2.hit_updates;
    last.node_visits_l0 += t20.node_visits; last.leaf_prim_tests_l0 += t20.prim_tests; last.queue_pushes_l0 += t20.pushes;
    last.staged_bytes_l0 += t20.staged_bytes; last.hit_updates_l0 += t20.hit_updates;
  }
  last.waves_retried = waves_retried; last.queues_grown = queues_grown; last.graph_replays = graph_replays;
  last.kernel_launches = launches + tracer.launches + tracer2.launches;
  last.traverse_launches = tracer.traverse_launches + tracer2.traverse_launches;
  last.traverse_launches_l0 = tracer.traverse_launches_l0 + tracer2.traverse_launches_l0;
  double l0a = 0, l0b = 0;
  last.ms_traverse = tracer.harvest_traverse_ms(&l0a) + tracer2.harvest_traverse_ms(&l0b);
  last.ms_traverse_l0 = l0a + l0b;
  last.ms_total = ms_total;
  return B2RT_OK;
}

int Renderer::stop() {
  if (!running) return B2RT_OK;
  RCHECK(set_device());
  // raise the cancel flag from the handle's cancel stream; the remaining waves generate no rays and are not accumulated
  k_fill_u32<<<1, 1, 0, stream_cancel>>>(counts + 3, 1u);
  B2RT_CUDA_OK(cudaStreamSynchronize(stream_cancel));
  return wait();
}

int Renderer::resolve(bool want_ldr) {
  RCHECK(set_device());
  if (running) RCHECK(wait());
  if (!accum) { set_error("no frame buffer"); return B2RT_ERR_INVALID; }
  const uint32_t np = width * height;
  k_resolve_image<<<(np + 255) / 256, 256, 0, stream>>>((const float4*)accum, (float4*)img_a, np, 0.f);
  resolved = img_a;
  if (cfg.median_threshold && samples_done < cfg.median_threshold) {
    dim3 grid((width + 31) / 32, (height + 7) / 8), block(32, 8);
    if (cfg.filter_kind == 1) {
      k_filter<1, false><<<grid, block, 0, stream>>>((const float4*)img_a, (float4*)img_b, width, height, 0.f);
    } else if (cfg.filter_kind == 2) {
      const float sr = cfg.filter_sigma_r > 0.f ? cfg.filter_sigma_r : 0.25f;
      k_filter<2, true><<<grid, block, 0, stream>>>((const float4*)img_a, (float4*)img_b, width, height, 1.0f / (sr * sr));
    } else {
      k_median3x3<<<grid, block, 0, stream>>>((const float4*)img_a, (float4*)img_b, width, height);
    }
    resolved = img_b;
  }
  if (want_ldr) k_tonemap<<<(np + 255) / 256, 256, 0, stream>>>((const float4*)resolved, ldr, np);
  B2RT_CUDA_OK(cudaGetLastError());
  return B2RT_OK;
}

void Renderer::fill_stats(b2rt_stats* out) const {
  *out = last;
  out->ms_build = build_ms;
  out->bvh_nodes = n_wide_nodes; out->bvh_subtrees = dbvh.n_treelets; out->bvh_levels = dbvh.n_levels;
  out->bvh_width = dbvh.width; out->bvh_bytes = dbvh.blob_bytes;
}

}  // namespace b2rt
