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
#include <cstring>

#include "render.cuh"
#include "rt_device.cuh"

namespace b2rt {

namespace {

constexpr float INF_F = __builtin_huge_valf();
constexpr uint32_t MAX_DEPTH = 96, ACT0 = 8, SH0 = 8 + MAX_DEPTH + 8, N_COUNTS = 256;

struct WaveParams {
  // wave geometry
  uint32_t pix0, n_pix;       // pixel range of this wave
  uint32_t spp;               // samples per pixel in this wave
  uint32_t sample0;           // global index of the wave's first sample
  uint32_t sample_stride;
  uint32_t width, height;
  uint32_t jitter;            // 0 -> pixel centre
  uint32_t k0, k1;            // Philox key
  float eps;
  uint32_t max_depth, ns_area_light, S;  // S = shadow rays per interaction
};

struct CamDev { f3 pos, cx, cy, cz; float tan_h, tan_v; };

struct SceneDev {
  const float4* prim_geom;      // 3 float4 per prim, scene order
  const float* tri_normals;     // 9 per tri or nullptr
  const uint32_t* prim_material;
  const b2rt_material* materials;
  const b2rt_light* lights;
  const float* light_area;
  uint32_t n_tris, n_lights;
};

struct PathBufs {
  float4* ray_o; float4* ray_d; unsigned long long* hits;
  float4* thr;     // rgb throughput, w = count_emission flag
  float4* rad;     // rgb radiance of the path
  float4* s_o; float4* s_d; unsigned long long* s_hits; float4* s_contrib;  // shadow rays [slot*S + j]
  uint32_t* ids_a; uint32_t* ids_b; uint32_t* s_ids;
  uint32_t* counts;  // [3] = cancel flag; [ACT0 + b] = active paths at bounce b; [SH0 + b] = shadow rays of bounce b
};

__global__ void __launch_bounds__(256)
k_raygen(WaveParams wp, CamDev cam, PathBufs pb) {
  const uint32_t n = wp.n_pix * wp.spp;
  const uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x;
  if (blockIdx.x == 0 && threadIdx.x < MAX_DEPTH + 1) {   // per-bounce list counters of this wave
    pb.counts[ACT0 + threadIdx.x] = (threadIdx.x == 0) ? (pb.counts[3] ? 0u : n) : 0u;
    pb.counts[SH0 + threadIdx.x] = 0u;
  }
  if (slot >= n) return;
  const uint32_t pix = wp.pix0 + slot / wp.spp;
  const uint32_t sample = wp.sample0 + (slot % wp.spp) * wp.sample_stride;
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
  pb.ray_o[slot] = make_float4(cam.pos.x, cam.pos.y, cam.pos.z, 0.0f);
  pb.ray_d[slot] = make_float4(d.x, d.y, d.z, INF_F);
  pb.hits[slot] = pack_hit(INF_F, 0xFFFFFFFFu);
  pb.thr[slot] = make_float4(1.f, 1.f, 1.f, 1.f);
  pb.rad[slot] = make_float4(0.f, 0.f, 0.f, 0.f);
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

__device__ __forceinline__ void append_id(uint32_t* list, uint32_t* counter, bool pred, uint32_t value) {
  const uint32_t m = __ballot_sync(__activemask(), pred);
  if (!pred) return;
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t leader = __ffs(m) - 1;
  uint32_t base = 0;
  if (lane == leader) base = atomicAdd(counter, (uint32_t)__popc(m));
  base = __shfl_sync(m, base, leader);
  list[base + __popc(m & ((1u << lane) - 1u))] = value;
}

// One surface interaction for every active path (bounce index b).
__global__ void __launch_bounds__(256)
k_shade(WaveParams wp, SceneDev sc, PathBufs pb, const uint32_t* __restrict__ ids, uint32_t* __restrict__ ids_next,
        uint32_t b, uint32_t identity_ids) {
  const uint32_t n = pb.counts[ACT0 + b];
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  bool cont = false;
  uint32_t slot = 0;
  if (i < n) {
    slot = identity_ids ? i : ids[i];
    const unsigned long long h = pb.hits[slot];
    const uint32_t prim = (uint32_t)h;
    const uint32_t S = wp.S;
    // default: no shadow rays
    for (uint32_t j = 0; j < S; ++j) pb.s_contrib[(size_t)slot * S + j].w = 0.f;
    if (prim != 0xFFFFFFFFu) {
      const float t = __uint_as_float((uint32_t)(h >> 32));
      const b2rt_material m = sc.materials[sc.prim_material[prim]];
      const float4 thr4 = pb.thr[slot];
      f3 thr = mk3(thr4.x, thr4.y, thr4.z);
      const bool count_emission = thr4.w != 0.f;
      if (m.kind == B2RT_MAT_EMISSION) {
        if (count_emission) {
          float4 L = pb.rad[slot];
          const f3 add = thr * mk3(m.emission[0], m.emission[1], m.emission[2]);
          L.x = L.x + add.x; L.y = L.y + add.y; L.z = L.z + add.z;
          pb.rad[slot] = L;
        }
      } else {
        const float4 ro = pb.ray_o[slot], rd = pb.ray_d[slot];
        const f3 o = mk3(ro.x, ro.y, ro.z), d = mk3(rd.x, rd.y, rd.z);
        const f3 P = o + d * t;
        PrimRec pr;
        pr.a = sc.prim_geom[(size_t)prim * 3]; pr.b = sc.prim_geom[(size_t)prim * 3 + 1]; pr.c = sc.prim_geom[(size_t)prim * 3 + 2];
        const bool is_sphere = prim >= sc.n_tris;
        f3 nrm;
        bool backface = false;
        if (is_sphere) {
          nrm = normalize3(P - mk3(pr.a.x, pr.a.y, pr.a.z));
        } else if (sc.tri_normals) {
          // recover (u,v) with the same arithmetic as the traversal test
          float tt, u = 0.f, v = 0.f;
          hit_triangle(pr, o, d, ro.w, INF_F, &tt, &u, &v);
          const float* nn = sc.tri_normals + (size_t)prim * 9;
          const float w0 = 1.0f - u - v;
          nrm = mk3(nn[3], nn[4], nn[5]) * u + mk3(nn[6], nn[7], nn[8]) * v + mk3(nn[0], nn[1], nn[2]) * w0;
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
        const uint32_t sample = wp.sample0 + (slot % wp.spp) * wp.sample_stride;

        if (m.kind == B2RT_MAT_DIFFUSE) {
          // direct lighting: src/pathtracer.cpp:439-478 + shadow ray (Task 4)
          uint32_t j = 0;
          for (uint32_t li = 0; li < sc.n_lights; ++li) {
            const b2rt_light lt = sc.lights[li];
            const uint32_t ns = lt.kind == B2RT_LIGHT_AREA ? max(1u, wp.ns_area_light) : 1u;
            for (uint32_t k = 0; k < ns; ++k, ++j) {
              f3 wi; float dist, pdf; f3 Lr;
              const f3 lp = mk3(lt.position[0], lt.position[1], lt.position[2]);
              const f3 ld = mk3(lt.direction[0], lt.direction[1], lt.direction[2]);
              const f3 radc = mk3(lt.radiance[0], lt.radiance[1], lt.radiance[2]);
              if (lt.kind == B2RT_LIGHT_AREA) {
                // AreaLight::sample_L, src/static_scene/light.cpp:81-92
                const uint4 r4 = philox4x32_10(pix, sample, b, 1 + j, wp.k0, wp.k1);
                const float ux = u01(r4.x) - 0.5f, uy = u01(r4.y) - 0.5f;
                const f3 dv = lp + mk3(lt.dim_x[0], lt.dim_x[1], lt.dim_x[2]) * ux + mk3(lt.dim_y[0], lt.dim_y[1], lt.dim_y[2]) * uy - P;
                const float sq = dot3(dv, dv);
                dist = __fsqrt_rn(sq);
                const float invd = __fdiv_rn(1.0f, dist);
                wi = dv * invd;
                const float cosT = dot3(wi, ld);
                pdf = __fdiv_rn(sq, sc.light_area[li] * fabsf(cosT));
                Lr = cosT < 0.0f ? radc : mk3(0, 0, 0);
              } else if (lt.kind == B2RT_LIGHT_POINT) {
                const f3 dv = lp - P;
                const float sq = dot3(dv, dv);
                dist = __fsqrt_rn(sq);
                wi = dv * __fdiv_rn(1.0f, dist);
                pdf = 1.0f; Lr = radc;
              } else {
                wi = neg3(ld); dist = INF_F; pdf = 1.0f; Lr = radc;
              }
              const float cos_in = dot3(wi, Z);
              if (!(cos_in >= 0.0f)) continue;
              if (!(Lr.x > 0.0f || Lr.y > 0.0f || Lr.z > 0.0f)) continue;
              if (!(pdf > 0.0f)) continue;
              const float wgt = __fdiv_rn(cos_in, (float)ns * pdf);
              const f3 f = mk3(m.albedo[0], m.albedo[1], m.albedo[2]) * 0.318309886183790672f;
              const f3 c = thr * f * Lr * wgt;
              const size_t sid = (size_t)slot * S + j;
              const float tmx = dist - wp.eps;
              pb.s_o[sid] = make_float4(P.x, P.y, P.z, wp.eps);
              pb.s_d[sid] = make_float4(wi.x, wi.y, wi.z, tmx);
              pb.s_hits[sid] = pack_hit(tmx, 0xFFFFFFFFu);
              pb.s_contrib[sid] = make_float4(c.x, c.y, c.z, 1.f);
            }
          }
        }
        if (b + 1 < wp.max_depth) {
          const uint4 r4 = philox4x32_10(pix, sample, b, 0, wp.k0, wp.k1);
          const float u2 = u01(r4.z), u3 = u01(r4.w);
          f3 wi_l, weight;
          bool delta = false;
          if (m.kind == B2RT_MAT_DIFFUSE) {
            const float r = __fsqrt_rn(u2);
            float s, c;
            sincos2pi(u3, &s, &c);
            const float zz = 1.0f - u2;
            wi_l = mk3(r * c, r * s, __fsqrt_rn(zz < 0.0f ? 0.0f : zz));
            weight = mk3(m.albedo[0], m.albedo[1], m.albedo[2]);
          } else if (m.kind == B2RT_MAT_MIRROR) {
            wi_l = mk3(-wo.x, -wo.y, wo.z);
            weight = mk3(m.albedo[0], m.albedo[1], m.albedo[2]);
            delta = true;
          } else {
            delta = true;
            const float eta = backface ? m.ior : __fdiv_rn(1.0f, m.ior);
            const float cos_i = wo.z;
            const float sin2_t = eta * eta * (1.0f - cos_i * cos_i);
            const bool tir = !(sin2_t < 1.0f);
            const float cos_t = tir ? 0.0f : __fsqrt_rn(1.0f - sin2_t);
            float Fr = 1.0f;
            if (!tir) {
              const float ni = backface ? m.ior : 1.0f, nt = backface ? 1.0f : m.ior;
              const float rs = __fdiv_rn(ni * cos_i - nt * cos_t, ni * cos_i + nt * cos_t);
              const float rp = __fdiv_rn(nt * cos_i - ni * cos_t, nt * cos_i + ni * cos_t);
              Fr = 0.5f * (rs * rs + rp * rp);
            }
            bool reflect;
            if (m.kind == B2RT_MAT_GLASS) reflect = tir || (u2 < Fr);
            else reflect = tir;
            if (reflect) {
              wi_l = mk3(-wo.x, -wo.y, wo.z);
              weight = m.kind == B2RT_MAT_GLASS ? mk3(m.albedo[0], m.albedo[1], m.albedo[2]) : mk3(1, 1, 1);
            } else {
              wi_l = mk3(-wo.x * eta, -wo.y * eta, -cos_t);
              weight = mk3(m.transmittance[0], m.transmittance[1], m.transmittance[2]);
            }
          }
          if (weight.x > 0.0f || weight.y > 0.0f || weight.z > 0.0f) {
            thr = thr * weight;
            const f3 nd = normalize3(X * wi_l.x + Y * wi_l.y + Z * wi_l.z);
            pb.ray_o[slot] = make_float4(P.x, P.y, P.z, wp.eps);
            pb.ray_d[slot] = make_float4(nd.x, nd.y, nd.z, INF_F);
            pb.hits[slot] = pack_hit(INF_F, 0xFFFFFFFFu);
            pb.thr[slot] = make_float4(thr.x, thr.y, thr.z, delta ? 1.f : 0.f);
            cont = true;
          }
        }
      }
    }
  }
  append_id(ids_next, &pb.counts[ACT0 + b + 1], cont, slot);
  // shadow-ray id list for the any-hit trace: every valid (slot, j); the loop is uniform across the warp
  for (uint32_t j = 0; j < wp.S; ++j) {
    bool valid = false;
    uint32_t sid = 0;
    if (i < n) { sid = slot * wp.S + j; valid = pb.s_contrib[sid].w != 0.f; }
    append_id(pb.s_ids, &pb.counts[SH0 + b], valid, sid);
  }
}

// add the unoccluded light samples in sample order (deterministic), then advance the lists
__global__ void __launch_bounds__(256)
k_resolve_shadow(WaveParams wp, PathBufs pb, const uint32_t* __restrict__ ids, uint32_t identity_ids, uint32_t b) {
  const uint32_t n = pb.counts[ACT0 + b];
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t slot = identity_ids ? i : ids[i];
  const uint32_t S = wp.S;
  float4 L = pb.rad[slot];
  bool any = false;
  for (uint32_t j = 0; j < S; ++j) {
    const size_t sid = (size_t)slot * S + j;
    const float4 c = pb.s_contrib[sid];
    if (c.w != 0.f && (uint32_t)pb.s_hits[sid] == 0xFFFFFFFFu) {
      L.x = L.x + c.x; L.y = L.y + c.y; L.z = L.z + c.z;
      any = true;
    }
  }
  if (any) pb.rad[slot] = L;
}

__global__ void k_wave_end(PathBufs pb, uint32_t max_depth, unsigned long long* totals) {
  // totals: [0] bounce rays, [1] shadow rays
  unsigned long long nb = 0, ns = 0;
  for (uint32_t b = 0; b < max_depth; ++b) { if (b) nb += pb.counts[ACT0 + b]; ns += pb.counts[SH0 + b]; }
  totals[0] += nb; totals[1] += ns;
}

// per-pixel accumulation in sample order (deterministic): accum.rgb += L_s, accum.w += 1
__global__ void __launch_bounds__(256)
k_accumulate(WaveParams wp, PathBufs pb, float4* __restrict__ accum) {
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
  stream = own_stream;
  B2RT_CUDA_OK(cudaEventCreate(&ev_start));
  B2RT_CUDA_OK(cudaEventCreate(&ev_done));
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
  void* ptrs[] = {ray_o, ray_d, hits, thr, rad, s_o, s_d, s_hits, s_contrib, ids_a, ids_b, s_ids, counts, totals};
  for (void* p : ptrs) free_ptr(p);
  ray_o = ray_d = thr = rad = s_o = s_d = s_contrib = nullptr; hits = s_hits = nullptr;
  ids_a = ids_b = s_ids = counts = nullptr; totals = nullptr;
  wave_cap = 0;
}

void Renderer::destroy() {
  if (stream) cudaStreamSynchronize(stream);
  release_wave();
  tracer.release();
  release_scene();
  free_ptr(accum); free_ptr(img_a); free_ptr(img_b); free_ptr(ldr);
  if (ev_start) cudaEventDestroy(ev_start);
  if (ev_done) cudaEventDestroy(ev_done);
  if (own_stream) cudaStreamDestroy(own_stream);
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
  HostScene hs;
  RCHECK(make_host_scene(d, &hs));
  WideBVH wb;
  RCHECK(build_wide_bvh(hs, cfg.max_leaf_size, cfg.bvh_width, cfg.treelet_bytes, &wb));
  RCHECK(upload_bvh(wb, &dbvh));   // grow-only device buffers: no cudaMalloc/cudaFree when the scene fits
  bvh_stale = true;                // wave buffers are kept; the tracer re-binds its (small) per-subtree arrays
  n_wide_nodes = wb.n_wide_nodes;
  build_ms = wb.build_ms;
  n_tris = hs.n_tris;
  n_lights = (uint32_t)hs.lights.size();
  const size_t np = std::max<size_t>(1, hs.n_prims());
  if (cap_prims < np) {
    free_ptr(d_prim_geom); free_ptr(d_prim_material); d_prim_geom = nullptr; d_prim_material = nullptr;
    cap_prims = np + np / 4;
    B2RT_CUDA_OK(cudaMalloc(&d_prim_geom, cap_prims * PRIM_BYTES));
    B2RT_CUDA_OK(cudaMalloc(&d_prim_material, cap_prims * 4));
  }
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
    B2RT_CUDA_OK(cudaMemcpy(d_prim_geom, hs.prim_geom.data(), (size_t)hs.n_prims() * PRIM_BYTES, cudaMemcpyHostToDevice));
    B2RT_CUDA_OK(cudaMemcpy(d_prim_material, hs.prim_material.data(), (size_t)hs.n_prims() * 4, cudaMemcpyHostToDevice));
  }
  B2RT_CUDA_OK(cudaMemcpy(d_materials, hs.materials.data(), hs.materials.size() * sizeof(b2rt_material), cudaMemcpyHostToDevice));
  if (!hs.tri_normals.empty()) {
    if (cap_normals < hs.tri_normals.size()) {
      free_ptr(d_tri_normals_buf); d_tri_normals_buf = nullptr;
      cap_normals = hs.tri_normals.size() + hs.tri_normals.size() / 4;
      B2RT_CUDA_OK(cudaMalloc(&d_tri_normals_buf, cap_normals * 4));
    }
    d_tri_normals = d_tri_normals_buf;
    B2RT_CUDA_OK(cudaMemcpy(d_tri_normals, hs.tri_normals.data(), hs.tri_normals.size() * 4, cudaMemcpyHostToDevice));
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
  have_scene = true;
  B2RT_CUDA_OK(cudaDeviceSynchronize());   // uploads above used the legacy stream; work runs on `stream`
  return B2RT_OK;
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
  shadow_per_hit = S;
  const uint64_t n_pix = (uint64_t)width * height;
  uint64_t cap = cfg.max_wave_paths ? cfg.max_wave_paths : (32u << 20);   // ~200 B of state per path: 6.4 GB of the 180 GB
  cap = std::max<uint64_t>(cap, 1024);
  const uint64_t want = std::min<uint64_t>(cap, n_pix * std::max(1u, cfg.ns_aa));
  const uint32_t Salloc = std::max(1u, S);
  if (wave_cap >= want && wave_S >= Salloc && tracer.max_rays >= want * Salloc) {
    if (bvh_stale) { RCHECK(tracer.init(dbvh, tracer.max_rays, 4)); bvh_stale = false; }
    return B2RT_OK;
  }
  release_wave();
  tracer.release();
  wave_cap = want; wave_S = Salloc;
  B2RT_CUDA_OK(cudaMalloc(&ray_o, wave_cap * sizeof(float4)));
  B2RT_CUDA_OK(cudaMalloc(&ray_d, wave_cap * sizeof(float4)));
  B2RT_CUDA_OK(cudaMalloc(&hits, wave_cap * 8));
  B2RT_CUDA_OK(cudaMalloc(&thr, wave_cap * sizeof(float4)));
  B2RT_CUDA_OK(cudaMalloc(&rad, wave_cap * sizeof(float4)));
  B2RT_CUDA_OK(cudaMalloc(&s_o, wave_cap * Salloc * sizeof(float4)));
  B2RT_CUDA_OK(cudaMalloc(&s_d, wave_cap * Salloc * sizeof(float4)));
  B2RT_CUDA_OK(cudaMalloc(&s_hits, wave_cap * Salloc * 8));
  B2RT_CUDA_OK(cudaMalloc(&s_contrib, wave_cap * Salloc * sizeof(float4)));
  B2RT_CUDA_OK(cudaMalloc(&ids_a, wave_cap * 4));
  B2RT_CUDA_OK(cudaMalloc(&ids_b, wave_cap * 4));
  B2RT_CUDA_OK(cudaMalloc(&s_ids, wave_cap * Salloc * 4));
  B2RT_CUDA_OK(cudaMalloc(&counts, N_COUNTS * 4));
  B2RT_CUDA_OK(cudaMalloc(&totals, 8 * 8));
  B2RT_CUDA_OK(cudaMemset(counts, 0, N_COUNTS * 4));
  B2RT_CUDA_OK(cudaMemset(totals, 0, 8 * 8));
  RCHECK(tracer.init(dbvh, wave_cap * Salloc, 4));
  return B2RT_OK;
}

int Renderer::start() {
  if (!have_scene || !have_camera || !accum) { set_error("start: scene, camera and frame size must be set first"); return B2RT_ERR_INVALID; }
  if (cfg.ns_aa == 0) { set_error("ns_aa must be >= 1"); return B2RT_ERR_INVALID; }
  RCHECK(set_device());
  if (running) RCHECK(wait());
  RCHECK(ensure_wave());
  const uint32_t S = shadow_per_hit;
  const uint32_t max_depth = std::min(MAX_DEPTH, std::max(1u, cfg.max_ray_depth));
  const uint32_t stride = cfg.sample_stride ? cfg.sample_stride : 1;
  const uint64_t n_pix = (uint64_t)width * height;

  CamDev cd;
  cd.pos = f3{cam.pos[0], cam.pos[1], cam.pos[2]};
  cd.cx = f3{cam.c2w[0], cam.c2w[1], cam.c2w[2]};
  cd.cy = f3{cam.c2w[3], cam.c2w[4], cam.c2w[5]};
  cd.cz = f3{cam.c2w[6], cam.c2w[7], cam.c2w[8]};
  cd.tan_h = tanf(cam.hfov_deg * 0.5f * 0.01745329251994329577f);
  cd.tan_v = tanf(cam.vfov_deg * 0.5f * 0.01745329251994329577f);

  SceneDev sd;
  sd.prim_geom = (const float4*)d_prim_geom; sd.tri_normals = d_tri_normals; sd.prim_material = d_prim_material;
  sd.materials = d_materials; sd.lights = d_lights; sd.light_area = d_light_area; sd.n_tris = n_tris; sd.n_lights = n_lights;

  PathBufs pb;
  pb.ray_o = (float4*)ray_o; pb.ray_d = (float4*)ray_d; pb.hits = hits; pb.thr = (float4*)thr; pb.rad = (float4*)rad;
  pb.s_o = (float4*)s_o; pb.s_d = (float4*)s_d; pb.s_hits = s_hits; pb.s_contrib = (float4*)s_contrib;
  pb.ids_a = ids_a; pb.ids_b = ids_b; pb.s_ids = s_ids; pb.counts = counts;

  tracer.launches = 0; tracer.traverse_launches = 0; tracer.ev_used = 0;
  launches = 0;
  B2RT_CUDA_OK(cudaMemsetAsync(totals, 0, 8 * 8, stream));
  B2RT_CUDA_OK(cudaMemsetAsync(tracer.counters, 0, sizeof(TraceCounters), stream));
  B2RT_CUDA_OK(cudaMemsetAsync(counts + 3, 0, 4, stream));  // cancel flag
  B2RT_CUDA_OK(cudaEventRecord(ev_start, stream));
  ms_traverse_acc = 0;

  // waves: pixel ranges x sample chunks, ascending in samples so accumulation order is fixed
  uint32_t spp_chunk, pix_chunk;
  if (n_pix >= wave_cap) { spp_chunk = 1; pix_chunk = (uint32_t)wave_cap; }
  else { spp_chunk = (uint32_t)std::min<uint64_t>(cfg.ns_aa, wave_cap / n_pix); pix_chunk = (uint32_t)n_pix; }
  cam_rays_enqueued = 0;
  for (uint32_t s0 = 0; s0 < cfg.ns_aa; s0 += spp_chunk) {
    const uint32_t spp = std::min(spp_chunk, cfg.ns_aa - s0);
    for (uint64_t p0 = 0; p0 < n_pix; p0 += pix_chunk) {
      WaveParams wp;
      wp.pix0 = (uint32_t)p0; wp.n_pix = (uint32_t)std::min<uint64_t>(pix_chunk, n_pix - p0);
      wp.spp = spp; wp.sample0 = cfg.sample_first + s0 * stride; wp.sample_stride = stride;
      wp.width = width; wp.height = height;
      wp.jitter = (cfg.ns_aa * stride) > 1 ? 1u : 0u;
      wp.k0 = (uint32_t)cfg.seed; wp.k1 = (uint32_t)(cfg.seed >> 32);
      wp.eps = cfg.ray_eps > 0.f ? cfg.ray_eps : 1e-4f;
      wp.max_depth = max_depth; wp.ns_area_light = cfg.ns_area_light; wp.S = S;
      const uint32_t n = wp.n_pix * wp.spp;
      const uint32_t g = (n + 255) / 256;
      k_raygen<<<g, 256, 0, stream>>>(wp, cd, pb); launches++;
      cam_rays_enqueued += n;
      uint32_t* cur = ids_a; uint32_t* nxt = ids_b;
      for (uint32_t b = 0; b < max_depth; ++b) {
        const uint32_t identity = b == 0 ? 1u : 0u;
        RCHECK(tracer.trace(stream, pb.ray_o, pb.ray_d, pb.hits, identity ? nullptr : cur, counts + ACT0 + b, false));
        k_shade<<<g, 256, 0, stream>>>(wp, sd, pb, cur, nxt, b, identity); launches++;
        if (S > 0) {
          RCHECK(tracer.trace(stream, pb.s_o, pb.s_d, pb.s_hits, pb.s_ids, counts + SH0 + b, true));
          k_resolve_shadow<<<g, 256, 0, stream>>>(wp, pb, cur, identity, b); launches++;
        }
        std::swap(cur, nxt);
      }
      k_wave_end<<<1, 1, 0, stream>>>(pb, max_depth, totals); launches++;
      k_accumulate<<<(wp.n_pix + 255) / 256, 256, 0, stream>>>(wp, pb, (float4*)accum); launches++;
    }
  }
  B2RT_CUDA_OK(cudaGetLastError());
  B2RT_CUDA_OK(cudaEventRecord(ev_done, stream));
  running = true;
  samples_pending = cfg.ns_aa;
  return B2RT_OK;
}

int Renderer::is_done() {
  if (!running) return 1;
  cudaError_t e = cudaEventQuery(ev_done);
  if (e == cudaSuccess) { int rc = wait(); return rc ? rc : 1; }
  if (e == cudaErrorNotReady) return 0;
  set_error(std::string("cudaEventQuery: ") + cudaGetErrorString(e));
  return B2RT_ERR_CUDA;
}

int Renderer::wait() {
  if (!running) return B2RT_OK;
  RCHECK(set_device());
  B2RT_CUDA_OK(cudaEventSynchronize(ev_done));
  running = false;
  samples_done += samples_pending;
  samples_pending = 0;
  float ms = 0;
  B2RT_CUDA_OK(cudaEventElapsedTime(&ms, ev_start, ev_done));
  ms_total = ms;
  unsigned long long t[8];
  B2RT_CUDA_OK(cudaMemcpy(t, totals, sizeof t, cudaMemcpyDeviceToHost));
  TraceCounters tc;
  B2RT_CUDA_OK(cudaMemcpy(&tc, tracer.counters, sizeof tc, cudaMemcpyDeviceToHost));
  uint32_t cancelled = 0;
  B2RT_CUDA_OK(cudaMemcpy(&cancelled, counts + 3, 4, cudaMemcpyDeviceToHost));
  last = b2rt_stats();
  last.rays_camera = cancelled ? 0 : cam_rays_enqueued;
  last.rays_bounce = t[0]; last.rays_shadow = t[1];
  last.node_visits = tc.node_visits; last.leaf_prim_tests = tc.prim_tests; last.subtree_visits = tc.subtree_visits;
  last.queue_pushes = tc.pushes; last.staged_bytes = tc.staged_bytes; last.hit_updates = tc.hit_updates;
  last.kernel_launches = launches + tracer.launches;
  last.traverse_launches = tracer.traverse_launches;
  last.ms_traverse = tracer.harvest_traverse_ms();
  last.ms_total = ms_total;
  bool ovf = false;
  RCHECK(tracer.check_overflow(stream, &ovf));
  if (ovf) { set_error("ray queue overflow: lower max_wave_paths"); return B2RT_ERR_OVERFLOW; }
  return B2RT_OK;
}

int Renderer::stop() {
  if (!running) return B2RT_OK;
  RCHECK(set_device());
  // raise the cancel flag from a second stream; remaining waves generate no rays
  cudaStream_t s2;
  B2RT_CUDA_OK(cudaStreamCreateWithFlags(&s2, cudaStreamNonBlocking));
  k_fill_u32<<<1, 1, 0, s2>>>(counts + 3, 1u);
  B2RT_CUDA_OK(cudaStreamSynchronize(s2));
  cudaStreamDestroy(s2);
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
    k_median3x3<<<grid, block, 0, stream>>>((const float4*)img_a, (float4*)img_b, width, height);
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
