// Device math of the b2rt kernels: fp32 vector ops, the canonical primitive tests, Philox RNG and
// the polynomial sincos.  The library is compiled with -fmad=false: the only fused multiply-adds
// are the explicit __fmaf_rn below, so results are bit-reproducible against any IEEE-754 CPU
// restatement of the same expressions (DESIGN.md "Arithmetic contract").
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace b2rt {

struct f3 { float x, y, z; };
__device__ __forceinline__ f3 mk3(float x, float y, float z) { f3 r; r.x = x; r.y = y; r.z = z; return r; }
__device__ __forceinline__ f3 operator+(f3 a, f3 b) { return mk3(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ f3 operator-(f3 a, f3 b) { return mk3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ f3 operator*(f3 a, float s) { return mk3(a.x * s, a.y * s, a.z * s); }
__device__ __forceinline__ f3 operator*(f3 a, f3 b) { return mk3(a.x * b.x, a.y * b.y, a.z * b.z); }
__device__ __forceinline__ f3 neg3(f3 a) { return mk3(-a.x, -a.y, -a.z); }
__device__ __forceinline__ float dot3(f3 a, f3 b) { return __fmaf_rn(a.z, b.z, __fmaf_rn(a.y, b.y, a.x * b.x)); }
__device__ __forceinline__ f3 cross3(f3 a, f3 b) {
  return mk3(__fmaf_rn(a.y, b.z, -(a.z * b.y)), __fmaf_rn(a.z, b.x, -(a.x * b.z)), __fmaf_rn(a.x, b.y, -(a.y * b.x)));
}
__device__ __forceinline__ f3 normalize3(f3 a) {
  float l = __fsqrt_rn(dot3(a, a));
  float inv = __fdiv_rn(1.0f, l);
  return a * inv;
}

// 48-byte primitive record, see b2rt_internal.h
struct PrimRec { float4 a, b, c; };

// Ray-triangle test in the reference's Moller-Trumbore form (src/static_scene/triangle.cpp:170-187:
// s = o - p1, t1 = e1 x d, t2 = s x e2, den = 1/dot(t1,e2), u = dot(-t2,d)*den, v = dot(t1,s)*den,
// t = dot(-t2,e1)*den, reject |den| > 1e10), fp32, accept set written positively (NaN rejects).
__device__ __forceinline__ bool hit_triangle(const PrimRec& p, f3 o, f3 d, float tmin, float tmax, float* t_out,
                                             float* u_out, float* v_out) {
  f3 v0 = mk3(p.a.x, p.a.y, p.a.z), e1 = mk3(p.a.w, p.b.x, p.b.y), e2 = mk3(p.b.z, p.b.w, p.c.x);
  f3 s = o - v0;
  f3 t1 = cross3(e1, d);
  f3 t2 = cross3(s, e2);
  float det = dot3(t1, e2);
  float den = __frcp_rn(det);
  if (!(fabsf(den) <= 1e10f)) return false;
  float u = -dot3(t2, d) * den;
  float v = dot3(t1, s) * den;
  float t = -dot3(t2, e1) * den;
  t = t + 0.0f;
  if (u >= 0.0f && v >= 0.0f && u <= 1.0f && v <= 1.0f && (u + v) <= 1.0f && t >= tmin && t <= tmax) {
    *t_out = t; *u_out = u; *v_out = v;
    return true;
  }
  return false;
}

// Ray-sphere (contract of Sphere::intersect, src/static_scene/sphere.h; body is a stub in the
// reference): nearest root inside [tmin,tmax].
__device__ __forceinline__ bool hit_sphere(const PrimRec& p, f3 o, f3 d, float tmin, float tmax, float* t_out) {
  f3 oc = o - mk3(p.a.x, p.a.y, p.a.z);
  float r = p.a.w;
  float a = dot3(d, d);
  float b = dot3(oc, d);
  float c = dot3(oc, oc) - r * r;
  float disc = __fmaf_rn(b, b, -(a * c));
  if (!(disc >= 0.0f)) return false;
  float sq = __fsqrt_rn(disc);
  float t1 = __fdiv_rn(-b - sq, a);
  float t2 = __fdiv_rn(-b + sq, a);
  t1 = t1 + 0.0f; t2 = t2 + 0.0f;
  if (t1 >= tmin && t1 <= tmax) { *t_out = t1; return true; }
  if (t2 >= tmin && t2 <= tmax) { *t_out = t2; return true; }
  return false;
}

__device__ __forceinline__ unsigned long long pack_hit(float t, uint32_t prim) {
  return ((unsigned long long)__float_as_uint(t) << 32) | prim;
}

// Philox4x32-10 (Salmon et al., SC'11): counter (c0..c3), key (k0,k1)
__device__ __forceinline__ uint4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint32_t hi0 = __umulhi(M0, c0), lo0 = M0 * c0, hi1 = __umulhi(M1, c2), lo1 = M1 * c2;
    uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  return make_uint4(c0, c1, c2, c3);
}
__device__ __forceinline__ float u01(uint32_t x) { return (float)(x >> 8) * (1.0f / 16777216.0f); }

// sin, cos of 2*pi*u: quadrant reduction + Taylor polynomials in plain mul/add
__device__ __forceinline__ void sincos2pi(float u, float* s_out, float* c_out) {
  float x = u * 4.0f;
  int k = (int)x;
  k = k > 3 ? 3 : (k < 0 ? 0 : k);
  float r = x - (float)k;
  float a = r * 1.57079632679489662f;
  float a2 = a * a;
  float ps = -1.0f / 6227020800.0f;
  ps = ps * a2 + 1.0f / 39916800.0f;
  ps = ps * a2 - 1.0f / 362880.0f;
  ps = ps * a2 + 1.0f / 5040.0f;
  ps = ps * a2 - 1.0f / 120.0f;
  ps = ps * a2 + 1.0f / 6.0f;
  ps = ps * a2;
  float s = a - a * ps;
  float pc = 1.0f / 479001600.0f;
  pc = pc * a2 - 1.0f / 3628800.0f;
  pc = pc * a2 + 1.0f / 40320.0f;
  pc = pc * a2 - 1.0f / 720.0f;
  pc = pc * a2 + 1.0f / 24.0f;
  pc = pc * a2 - 0.5f;
  float c = 1.0f + pc * a2;
  float so, co;
  if (k == 0) { so = s; co = c; }
  else if (k == 1) { so = c; co = -s; }
  else if (k == 2) { so = -s; co = -c; }
  else { so = -c; co = s; }
  *s_out = so; *c_out = co;
}

// atan2(y, x) / (2 pi) in [0, 1): octant reduction + the polynomial of Abramowitz & Stegun 4.4.49 in plain fp32
// multiplies / adds and one IEEE division -- the same expression as the oracle's, so environment-map look-ups are
// bit-identical (libm's atan2f / acosf differ between glibc and CUDA)
__device__ __forceinline__ float atan2_turns(float y, float x) {
  const float ax = fabsf(x), ay = fabsf(y);
  const bool swap = ay > ax;
  const float num = swap ? ax : ay, den = swap ? ay : ax;
  const float a = den > 0.0f ? __fdiv_rn(num, den) : 0.0f;
  const float a2 = a * a;
  float p = 0.0028662257f;
  p = p * a2 - 0.0161657367f;
  p = p * a2 + 0.0429096138f;
  p = p * a2 - 0.0752896400f;
  p = p * a2 + 0.1065626393f;
  p = p * a2 - 0.1420889944f;
  p = p * a2 + 0.1999355085f;
  p = p * a2 - 0.3333314528f;
  p = p * a2 + 1.0f;
  float r = (a * p) * 0.159154943091895336f;
  if (swap) r = 0.25f - r;
  if (x < 0.0f) r = 0.5f - r;
  if (y < 0.0f) r = 1.0f - r;
  return r >= 1.0f ? 0.0f : r;
}

__device__ __forceinline__ float powi(float b, uint32_t e) {
  float r = 1.0f;
  while (e) { if (e & 1u) r = r * b; b = b * b; e >>= 1; }
  return r;
}
// exponent of the glossy lobe (B2RT_MAT_GLOSSY, include/b2rt.h)
__device__ __forceinline__ uint32_t glossy_exponent(float roughness) {
  const float r2 = roughness * roughness;
  if (!(r2 > 4.8e-4f)) return 4096u;
  const float e = __fdiv_rn(2.0f, r2) - 2.0f;
  return e < 1.0f ? 1u : (e > 4096.0f ? 4096u : (uint32_t)e);
}

}  // namespace b2rt
