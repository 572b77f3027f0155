// ORACLE -- TEST INFRASTRUCTURE ONLY.  Not part of the product path.
//
// CPU restatement of the reference's path for parity checking and as the timed CPU baseline.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
// load this library.  The product (cuda-raytracer_b200/csrc) never links or calls it.
//
// PARITY PIN STATUS: the reference ships NO numeric golden vectors and its CPU traversal /
// integrator bodies are stubs (src/bvh.cpp:412-439, src/pathtracer.cpp:415-496, src/bsdf.cpp:41-96,
// src/camera.cpp:111-117, src/bbox.cpp:10-17).  What IS pinned against the real reference:
//   * the BVH builder + 4-wide collapse, bit-for-bit, against the reference's own src/bvh.cpp
//     compiled from where it lies (oracle/build_ref.sh -> oracle/_ref/ref_bvh_dump): primitive
//     order, every node's [start,range), wide node count and level profile;
//   * closest hit: the BVH traversal here is checked against an exhaustive argmin over all
//     primitives using the reference's ray-triangle arithmetic (triangle.cpp:170-209).
// Closest-hit primitive ids and radiance are therefore "parity unpinned by the reference,
// pinned by this oracle" (see DESIGN.md).
//
// Every function cites the reference file:line it follows (paths relative to the reference).
// Arithmetic contract shared with the CUDA kernels (written independently there): fp32, no
// implicit contraction (-ffp-contract=off / -fmad=false), explicit fmaf where stated below.

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <limits>
#include <thread>
#include <vector>

#include "../include/b2rt.h"  // interface structs only (scene desc, camera, config)

namespace {

struct V3 { float x, y, z; };
inline V3 mk(float x, float y, float z) { return V3{x, y, z}; }
inline V3 operator+(V3 a, V3 b) { return mk(a.x + b.x, a.y + b.y, a.z + b.z); }
inline V3 operator-(V3 a, V3 b) { return mk(a.x - b.x, a.y - b.y, a.z - b.z); }
inline V3 operator*(V3 a, float s) { return mk(a.x * s, a.y * s, a.z * s); }
inline V3 operator*(V3 a, V3 b) { return mk(a.x * b.x, a.y * b.y, a.z * b.z); }
inline V3 neg(V3 a) { return mk(-a.x, -a.y, -a.z); }
// contract: dot = fma(z,z, fma(y,y, x*x)); cross_i = fma(a_j, b_k, -(a_k*b_j))
inline float dot(V3 a, V3 b) { return fmaf(a.z, b.z, fmaf(a.y, b.y, a.x * b.x)); }
inline V3 cross(V3 a, V3 b) {
  return mk(fmaf(a.y, b.z, -(a.z * b.y)), fmaf(a.z, b.x, -(a.x * b.z)), fmaf(a.x, b.y, -(a.y * b.x)));
}
inline V3 normalize(V3 a) {
  float l = sqrtf(dot(a, a));
  float inv = 1.0f / l;
  return a * inv;
}

// ---- Philox4x32-10 (Salmon et al. 2011), counter-based RNG keyed per (pixel, sample, bounce) ----
// Replaces cuRAND XORWOW states (src/samplers.cu_inl:8-40, src/cudaRenderer.cu:1299-1302) and
// std::rand (src/sampler.cpp:17-18).  Not pinned by the reference by design.
inline void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
                          uint32_t out[4]) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
  for (int r = 0; r < 10; ++r) {
    uint64_t p0 = (uint64_t)M0 * c0, p1 = (uint64_t)M1 * c2;
    uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0, hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
    uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += W0; k1 += W1;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
inline float u01(uint32_t x) { return (float)(x >> 8) * (1.0f / 16777216.0f); }

// sin/cos of 2*pi*u by quadrant reduction + Taylor polynomials, plain mul/add (bit-reproducible
// on CPU and GPU; libm / CUDA sinf differ in the last ulp).
inline void sincos2pi(float u, float* s_out, float* c_out) {
  float x = u * 4.0f;
  int k = (int)x;
  if (k > 3) k = 3;
  if (k < 0) k = 0;
  float r = x - (float)k;
  float a = r * 1.57079632679489662f;
  float a2 = a * a;
  float ps = -1.0f / 6227020800.0f;          // a^13
  ps = ps * a2 + 1.0f / 39916800.0f;         // a^11
  ps = ps * a2 - 1.0f / 362880.0f;           // a^9
  ps = ps * a2 + 1.0f / 5040.0f;
  ps = ps * a2 - 1.0f / 120.0f;
  ps = ps * a2 + 1.0f / 6.0f;
  ps = ps * a2;                              // a2*(1/6 - ...)
  float s = a - a * ps;
  float pc = 1.0f / 479001600.0f;            // a^12
  pc = pc * a2 - 1.0f / 3628800.0f;          // a^10
  pc = pc * a2 + 1.0f / 40320.0f;
  pc = pc * a2 - 1.0f / 720.0f;
  pc = pc * a2 + 1.0f / 24.0f;
  pc = pc * a2 - 0.5f;
  float c = 1.0f + pc * a2;
  switch (k) {
    case 0: *s_out = s; *c_out = c; break;
    case 1: *s_out = c; *c_out = -s; break;
    case 2: *s_out = -s; *c_out = -c; break;
    default: *s_out = -c; *c_out = s; break;
  }
}

// atan2(y, x) / (2 pi) in [0, 1): octant reduction + the polynomial of Abramowitz & Stegun 4.4.49 (|err| <= 2e-8 on
// [0, 1]) in plain fp32 multiplies / adds and one IEEE division, restated identically in the kernels (rt_device.cuh),
// so that environment-map lookups are bit-identical (libm's atan2f / acosf differ between glibc and CUDA).
inline float atan2_turns(float y, float x) {
  const float ax = fabsf(x), ay = fabsf(y);
  const bool swap = ay > ax;
  const float num = swap ? ax : ay, den = swap ? ay : ax;
  const float a = den > 0.0f ? num / den : 0.0f;
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
  float r = (a * p) * 0.159154943091895336f;   // turns
  if (swap) r = 0.25f - r;
  if (x < 0.0f) r = 0.5f - r;
  if (y < 0.0f) r = 1.0f - r;
  return r >= 1.0f ? 0.0f : r;
}

// integer power by squaring (the glossy lobe's exponent is an integer so that no pow() is involved)
inline float powi(float b, uint32_t e) {
  float r = 1.0f;
  while (e) { if (e & 1u) r = r * b; b = b * b; e >>= 1; }
  return r;
}
// GlossyBSDF(reflectance, roughness) (src/bsdf.h:143-162, commented out in the checkout; bodies are stubs,
// src/bsdf.cpp:59-70): a normalised Phong lobe about the mirror direction, exponent n = clamp(int(2 / roughness^2) - 2, 1, 4096)
inline uint32_t glossy_exponent(float roughness) {
  const float r2 = roughness * roughness;
  if (!(r2 > 4.8e-4f)) return 4096u;
  const float e = 2.0f / r2 - 2.0f;
  return e < 1.0f ? 1u : (e > 4096.0f ? 4096u : (uint32_t)e);
}

// ---- primitives -----------------------------------------------------------------------------------
struct Prim {  // triangle: v0,e1,e2 ; sphere: v0 = centre, e1.x = radius
  V3 v0, e1, e2;
  uint32_t id;
  uint32_t is_sphere;
};

struct Counters {
  uint64_t box_tests = 0, prim_tests = 0, rays_camera = 0, rays_bounce = 0, rays_shadow = 0;
  void add(const Counters& o) {
    box_tests += o.box_tests; prim_tests += o.prim_tests; rays_camera += o.rays_camera;
    rays_bounce += o.rays_bounce; rays_shadow += o.rays_shadow;
  }
};

// Ray-triangle: Moller-Trumbore in the reference's own form, src/static_scene/triangle.cpp:170-187
// (s, e1, e2, t1 = e1 x d, t2 = s x e2, den = 1/dot(t1,e2), u = dot(-t2,d)*den, v = dot(t1,s)*den,
//  t = dot(-t2,e1)*den; reject |den| > 1e10, u,v outside [0,1], u+v > 1, t outside [min_t,max_t]),
// restated in fp32.  Accept set written positively so NaNs reject.
inline bool hit_triangle(const Prim& p, V3 o, V3 d, float tmin, float tmax, float* t_out, float* u_out,
                         float* v_out) {
  V3 s = o - p.v0;
  V3 t1 = cross(p.e1, d);
  V3 t2 = cross(s, p.e2);
  float det = dot(t1, p.e2);
  float den = 1.0f / det;
  if (!(fabsf(den) <= 1e10f)) return false;
  float u = -dot(t2, d) * den;
  float v = dot(t1, s) * den;
  float t = -dot(t2, p.e1) * den;
  t = t + 0.0f;  // canonicalise -0
  if (u >= 0.0f && v >= 0.0f && u <= 1.0f && v <= 1.0f && (u + v) <= 1.0f && t >= tmin && t <= tmax) {
    *t_out = t; *u_out = u; *v_out = v;
    return true;
  }
  return false;
}

// Ray-sphere: contract of Sphere::intersect / Sphere::test (src/static_scene/sphere.cpp:11-36 is a
// stub; header src/static_scene/sphere.h documents t1 <= t2, nearest root inside [min_t,max_t]).
inline bool hit_sphere(const Prim& p, V3 o, V3 d, float tmin, float tmax, float* t_out) {
  V3 oc = o - p.v0;
  float r = p.e1.x;
  float a = dot(d, d);
  float b = dot(oc, d);
  float c = dot(oc, oc) - r * r;
  float disc = fmaf(b, b, -(a * c));
  if (!(disc >= 0.0f)) return false;
  float sq = sqrtf(disc);
  float t1 = (-b - sq) / a;
  float t2 = (-b + sq) / a;
  t1 = t1 + 0.0f; t2 = t2 + 0.0f;
  if (t1 >= tmin && t1 <= tmax) { *t_out = t1; return true; }
  if (t2 >= tmin && t2 <= tmax) { *t_out = t2; return true; }
  return false;
}

struct Hit {
  float t = std::numeric_limits<float>::infinity();
  uint32_t prim = 0xFFFFFFFFu;
  float u = 0, v = 0;
};

inline void test_prim(const Prim& p, V3 o, V3 d, float tmin, float tmax, Hit* best, Counters* cn) {
  cn->prim_tests++;
  float t, u = 0, v = 0;
  bool h = p.is_sphere ? hit_sphere(p, o, d, tmin, tmax, &t) : hit_triangle(p, o, d, tmin, tmax, &t, &u, &v);
  if (h && (t < best->t || (t == best->t && p.id < best->prim))) {
    best->t = t; best->prim = p.id; best->u = u; best->v = v;
  }
}

// ---- binary SAH BVH, restated from src/bvh.cpp:18-230 + src/bvh.cpp:339-365 --------------------------
struct BBoxD {
  double mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
  void expand(const BBoxD& b) {
    for (int i = 0; i < 3; ++i) { mn[i] = std::min(mn[i], b.mn[i]); mx[i] = std::max(mx[i], b.mx[i]); }
  }
  bool empty() const { return mn[0] > mx[0] || mn[1] > mx[1] || mn[2] > mx[2]; }
  double surface_area() const {  // src/bbox.h:108-112
    if (empty()) return 0.0;
    double ex = mx[0] - mn[0], ey = mx[1] - mn[1], ez = mx[2] - mn[2];
    return 2 * (ex * ez + ex * ey + ey * ez);
  }
  double centroid(int a) const { return (mn[a] + mx[a]) / 2; }
};

struct BNode {
  BBoxD bb;
  size_t start, range;
  int l = -1, r = -1;
};

struct Scene {
  std::vector<Prim> prims;          // scene order (prim id = index)
  std::vector<BBoxD> pbox;          // Triangle::get_bbox with PADDING 1e-3, triangle.cpp:13-47
  std::vector<float> normals;       // n_tris*9 or empty
  std::vector<uint32_t> prim_mat;
  std::vector<b2rt_material> mats;
  std::vector<b2rt_light> lights;
  uint32_t n_tris = 0, n_spheres = 0;
  // environment map (EnvironmentLight, src/static_scene/environment_light.h; PathTracer ctor argument envmap,
  // src/pathtracer.h:57-60): RGB fp32, index x + y*w, row 0 = the +y pole, x = azimuth atan2(z, x) / 2 pi
  std::vector<float> env; uint32_t env_w = 0, env_h = 0;
  // BVH
  std::vector<uint32_t> order;      // BVHAccel::primitives after the build (getSortedPrimitives)
  std::vector<BNode> nodes;         // pre-order, node 0 = root
  size_t max_leaf = 0;
};

struct Builder {
  Scene& sc;
  size_t max_leaf;
  explicit Builder(Scene& s, size_t ml) : sc(s), max_leaf(ml) {}
  double cen(uint32_t p, int a) const { return sc.pbox[p].centroid(a); }

  // splitBVHNode, src/bvh.cpp:48-230.  Same std::sort calls in the same order with the same
  // comparators (centroid of the padded box), 12 planes per axis, cost 5 + 2*SA-weighted counts.
  int split(size_t start, size_t end, const BBoxD& bb) {
    int idx = (int)sc.nodes.size();
    sc.nodes.push_back(BNode());
    sc.nodes[idx].bb = bb; sc.nodes[idx].start = start; sc.nodes[idx].range = end - start;
    if (end - start <= max_leaf) return idx;
    double total_sa = bb.surface_area();
    if (total_sa < 1e-15) return idx;

    auto& P = sc.order;
    float current_cost = 2 * (float)(end - start);
    float bestcost = current_cost;
    int besti = 0;
    float bestk = 0;
    BBoxD boxl, boxr;
    const int numparts = 12;
    for (int i = 0; i < 3; ++i) {
      std::sort(P.begin() + start, P.begin() + end, [&](uint32_t a, uint32_t b) { return cen(a, i) < cen(b, i); });
      std::vector<BBoxD> ltor, rtol;
      BBoxD b1, b2;
      double startval = cen(P[start], i), endval = cen(P[end - 1], i);
      int lastidx = (int)start;
      std::vector<int> indices;
      for (long part = 1; part <= numparts; ++part) {
        double divider = startval + part * ((endval - startval) / (numparts + 1));
        int id = (int)(std::upper_bound(P.begin() + start, P.begin() + end, divider,
                                        [&](double dv, uint32_t b) { return dv < cen(b, i); }) - P.begin());
        for (int j = lastidx; j < id; ++j) b1.expand(sc.pbox[P[j]]);
        indices.push_back(id);
        lastidx = id;
        ltor.push_back(b1);
      }
      lastidx = (int)end;
      for (long part = 1; part <= numparts; ++part) {
        double divider = endval - part * ((endval - startval) / (numparts + 1));
        int id = (int)(std::lower_bound(P.begin() + start, P.begin() + end, divider,
                                        [&](uint32_t b, double dv) { return cen(b, i) < dv; }) - P.begin());
        for (int j = lastidx - 1; j >= id; --j) b2.expand(sc.pbox[P[j]]);
        lastidx = id;
        rtol.push_back(b2);
      }
      double mincost = current_cost;
      size_t mink = 1;
      BBoxD minboxl, minboxr;
      for (size_t k = 0; k < (size_t)numparts; ++k) {
        int count = indices[k] - (int)start;
        int count2 = (int)(end - start) - count;
        double sa1 = ltor[k].surface_area();
        double sa2 = rtol[numparts - k - 1].surface_area();
        double cost = 5 + (sa1 / total_sa) * count * 2 + (sa2 / total_sa) * count2 * 2;
        if (mincost > cost) {
          mincost = cost; mink = indices[k]; minboxl = ltor[k]; minboxr = rtol[numparts - k - 1];
        }
      }
      if (mincost == current_cost) {  // bvh.cpp:193-197
        mink = indices[1]; minboxl = ltor[1]; minboxr = rtol[numparts - 2];
      }
      if (mincost < bestcost) {
        bestcost = (float)mincost; bestk = (float)mink; besti = i; boxl = minboxl; boxr = minboxr;
      }
    }
    if (bestcost == current_cost) return idx;
    std::sort(P.begin() + start, P.begin() + end,
              [&](uint32_t a, uint32_t b) { return cen(a, besti) < cen(b, besti); });
    size_t k = (size_t)bestk;  // the reference stores the split index in a float (bvh.cpp:62,202)
    int l = split(start, k, boxl);
    int r = split(k, end, boxr);
    sc.nodes[idx].l = l; sc.nodes[idx].r = r;
    return idx;
  }
};

void build_bvh(Scene& sc, size_t max_leaf) {
  sc.max_leaf = max_leaf;
  sc.nodes.clear();
  sc.order.resize(sc.prims.size());
  for (size_t i = 0; i < sc.order.size(); ++i) sc.order[i] = (uint32_t)i;
  if (sc.prims.empty()) return;
  BBoxD bb;
  for (auto& b : sc.pbox) bb.expand(b);
  Builder B(sc, max_leaf);
  // BVHAccel::BVHAccel sorts by x first, src/bvh.cpp:357
  std::sort(sc.order.begin(), sc.order.end(), [&](uint32_t a, uint32_t b) { return B.cen(a, 0) < B.cen(b, 0); });
  B.split(0, sc.order.size(), bb);
}

// 4-wide collapse statistics, restated from BVHNode::compactTree (src/bvh.cpp:275-337, DEPTH 2,
// TREE_BRANCHES 4) and BVHSubTree::compress (:234-273): wide node count per level.
void wide_levels(const Scene& sc, int node, int depth, std::vector<uint32_t>& levels) {
  if ((int)levels.size() <= depth) levels.resize(depth + 1, 0);
  levels[depth]++;
  const BNode& n = sc.nodes[node];
  if (n.l < 0 && n.r < 0) return;
  std::vector<std::pair<int, int>> st;
  st.push_back({0, node});
  while (!st.empty()) {
    auto dn = st.back(); st.pop_back();
    const BNode& m = sc.nodes[dn.second];
    if (dn.first == 2) { wide_levels(sc, dn.second, depth + 1, levels); continue; }
    if (m.l >= 0) st.push_back({dn.first + 1, m.l});
    if (m.r >= 0) st.push_back({dn.first + 1, m.r});
    if (m.l < 0 && m.r < 0) wide_levels(sc, dn.second, depth + 1, levels);
  }
}

// Ray-box: contract of BBox::intersect(r, t0, t1) (src/bbox.h:117-124; body is a stub in
// src/bbox.cpp:10-17).  Slab test in double on the padded boxes, conservative.
inline bool hit_box(const BBoxD& b, const double o[3], const double inv[3], double tmin, double tmax, double* tn) {
  double t0 = tmin, t1 = tmax;
  for (int a = 0; a < 3; ++a) {
    double ta = (b.mn[a] - o[a]) * inv[a], tb = (b.mx[a] - o[a]) * inv[a];
    if (ta > tb) std::swap(ta, tb);
    if (ta != ta || tb != tb) continue;  // 0 * inf: ray lies in the slab plane; padded boxes make this safe
    if (ta > t0) t0 = ta;
    if (tb < t1) t1 = tb;
  }
  *tn = t0;
  return t0 <= t1 * 1.0000001;
}

// Closest hit: front-to-back binary traversal, semantics of nodeIntersect in the stale
// src/static_scene/bvh.cpp:413-497 (nearest child first, far child skipped when the near hit is
// closer than the far box), leaf = nearest of its primitives.
void closest_bvh(const Scene& sc, V3 o, V3 d, float tmin, float tmax, Hit* best, Counters* cn) {
  if (sc.nodes.empty()) return;
  double od[3] = {o.x, o.y, o.z}, inv[3] = {1.0 / (double)d.x, 1.0 / (double)d.y, 1.0 / (double)d.z};
  struct E { int node; double tn; };
  E stack[128];
  int sp = 0;
  double tn;
  cn->box_tests++;
  if (!hit_box(sc.nodes[0].bb, od, inv, tmin, tmax, &tn)) return;
  stack[sp++] = {0, tn};
  while (sp) {
    E e = stack[--sp];
    if (e.tn > (double)best->t) continue;
    const BNode& n = sc.nodes[e.node];
    if (n.l < 0) {
      for (size_t p = 0; p < n.range; ++p) test_prim(sc.prims[sc.order[n.start + p]], o, d, tmin, tmax, best, cn);
      continue;
    }
    double tl, tr;
    cn->box_tests += 2;
    bool hl = hit_box(sc.nodes[n.l].bb, od, inv, tmin, std::min((double)tmax, (double)best->t), &tl);
    bool hr = hit_box(sc.nodes[n.r].bb, od, inv, tmin, std::min((double)tmax, (double)best->t), &tr);
    if (hl && hr) {
      if (tl < tr) { stack[sp++] = {n.r, tr}; stack[sp++] = {n.l, tl}; }
      else { stack[sp++] = {n.l, tl}; stack[sp++] = {n.r, tr}; }
    } else if (hl) stack[sp++] = {n.l, tl};
    else if (hr) stack[sp++] = {n.r, tr};
  }
}

void closest_brute(const Scene& sc, V3 o, V3 d, float tmin, float tmax, Hit* best, Counters* cn) {
  for (const Prim& p : sc.prims) test_prim(p, o, d, tmin, tmax, best, cn);
}

bool occluded_bvh(const Scene& sc, V3 o, V3 d, float tmin, float tmax, Counters* cn) {
  Hit h;
  closest_bvh(sc, o, d, tmin, tmax, &h, cn);  // any hit inside [tmin,tmax] <=> a closest hit exists
  return h.prim != 0xFFFFFFFFu;
}

// ---- shading -----------------------------------------------------------------------------------------
// make_coord_space, src/bsdf.cpp:14-33 (fp32)
inline void make_coord_space(V3 n, V3* X, V3* Y, V3* Z) {
  V3 z = n, h = z;
  if (fabsf(h.x) <= fabsf(h.y) && fabsf(h.x) <= fabsf(h.z)) h.x = 1.0f;
  else if (fabsf(h.y) <= fabsf(h.x) && fabsf(h.y) <= fabsf(h.z)) h.y = 1.0f;
  else h.z = 1.0f;
  z = normalize(z);
  V3 y = normalize(cross(h, z));
  V3 x = normalize(cross(z, y));
  *X = x; *Y = y; *Z = z;
}

struct Cam {
  V3 pos, cx, cy, cz;
  float tan_h, tan_v;
};

// Camera::generate_ray contract, src/camera.h:71-81 (body stub src/camera.cpp:111-117): (x,y) in
// [0,1]^2 on the sensor plane one unit in front of the pinhole, c2w from compute_position (:87-109).
inline void generate_ray(const Cam& c, float sx, float sy, V3* o, V3* d) {
  float px = (2.0f * sx - 1.0f) * c.tan_h;
  float py = (2.0f * sy - 1.0f) * c.tan_v;
  V3 w = c.cx * px + c.cy * py - c.cz;
  *o = c.pos;
  *d = normalize(w);
}

// bilinear look-up of the environment map in direction d (unit): wraps in azimuth, clamps at the poles
inline V3 env_lookup(const Scene& sc, V3 d) {
  const float u = atan2_turns(d.z, d.x);
  float cy = d.y > 1.0f ? 1.0f : (d.y < -1.0f ? -1.0f : d.y);
  float sy = 1.0f - cy * cy;
  const float v = 2.0f * atan2_turns(sqrtf(sy < 0.0f ? 0.0f : sy), cy);   // polar angle / pi
  const float fx = u * (float)sc.env_w - 0.5f, fy = v * (float)sc.env_h - 0.5f;
  const float flx = floorf(fx), fly = floorf(fy);
  const float tx = fx - flx, ty = fy - fly;
  int x0 = (int)flx, y0 = (int)fly;
  int x1 = x0 + 1, y1 = y0 + 1;
  const int W = (int)sc.env_w, H = (int)sc.env_h;
  x0 = ((x0 % W) + W) % W; x1 = ((x1 % W) + W) % W;
  y0 = y0 < 0 ? 0 : (y0 > H - 1 ? H - 1 : y0); y1 = y1 < 0 ? 0 : (y1 > H - 1 ? H - 1 : y1);
  auto px = [&](int x, int y) { const float* p = &sc.env[3 * ((size_t)y * W + x)]; return mk(p[0], p[1], p[2]); };
  const V3 a = px(x0, y0) * (1.0f - tx) + px(x1, y0) * tx;
  const V3 b = px(x0, y1) * (1.0f - tx) + px(x1, y1) * tx;
  return a * (1.0f - ty) + b * ty;
}

// f(wo, wi) of the non-delta BSDFs in the local frame (z = normal): diffuse albedo / pi (src/bsdf.cpp:37-39); glossy
// reflectance * (n + 2) / (2 pi) * cos^n(angle between wi and the mirror direction of wo)
inline V3 bsdf_f(const b2rt_material& m, V3 wo, V3 wi) {
  const V3 alb = mk(m.albedo[0], m.albedo[1], m.albedo[2]);
  if (m.kind == B2RT_MAT_GLOSSY) {
    const uint32_t n = glossy_exponent(m.roughness);
    float c = dot(mk(-wo.x, -wo.y, wo.z), wi);
    if (!(c > 0.0f)) c = 0.0f;
    return alb * (((float)(n + 2u) * 0.159154943091895336f) * powi(c, n));
  }
  return alb * 0.318309886183790672f;
}

struct RenderCtx {
  const Scene* sc;
  Cam cam;
  b2rt_config cfg;
  uint32_t w, h;
  float eps;
  uint32_t spp_total_for_jitter;
  std::vector<float> light_area;
};

// One path.  Estimator: src/pathtracer.cpp:395-496 skeleton (Le at the hit, direct lighting loop
// :439-478 with cos/(n*pdf)*f*L and the w_in.z < 0 skip), completed with shadow rays (Task 4),
// one BSDF-sampled indirect bounce per interaction (Task 5) and no Russian roulette.  Emitted
// radiance is counted for camera rays and after delta bounces only (next-event estimation
// covers the rest).  AreaLight::sample_L: src/static_scene/light.cpp:81-92, with the cosine
// normalised by the distance (the checkout's :88 omits it; see DESIGN.md "stated divergences").
V3 trace_path(const RenderCtx& rc, uint32_t pix, uint32_t x, uint32_t y, uint32_t sample, Counters* cn) {
  const Scene& sc = *rc.sc;
  const uint32_t k0 = (uint32_t)rc.cfg.seed, k1 = (uint32_t)(rc.cfg.seed >> 32);
  uint32_t r4[4];
  philox4x32_10(pix, sample, 0, 0, k0, k1, r4);
  float jx = 0.5f, jy = 0.5f;
  if (rc.spp_total_for_jitter > 1) { jx = u01(r4[0]); jy = u01(r4[1]); }
  float sx = ((float)x + jx) / (float)rc.w, sy = ((float)y + jy) / (float)rc.h;
  V3 o, d;
  generate_ray(rc.cam, sx, sy, &o, &d);
  float tmin = 0.0f;
  V3 thr = mk(1, 1, 1), L = mk(0, 0, 0);
  bool count_emission = true;
  const float INF = std::numeric_limits<float>::infinity();
  const uint32_t max_depth = rc.cfg.max_ray_depth < 1 ? 1 : rc.cfg.max_ray_depth;
  for (uint32_t b = 0; b < max_depth; ++b) {
    Hit hit;
    if (b == 0) cn->rays_camera++; else cn->rays_bounce++;
    closest_bvh(sc, o, d, tmin, INF, &hit, cn);
    if (hit.prim == 0xFFFFFFFFu) {
      // EnvironmentLight::sample_dir (src/static_scene/environment_light.h): counted like emitted radiance
      if (sc.env_w && count_emission) L = L + thr * env_lookup(sc, d);
      break;
    }
    const b2rt_material& m = sc.mats[sc.prim_mat[hit.prim]];
    if (m.kind == B2RT_MAT_EMISSION) {
      if (count_emission) L = L + thr * mk(m.emission[0], m.emission[1], m.emission[2]);
      break;
    }
    V3 P = o + d * hit.t;
    // shading normal: Triangle::intersect, triangle.cpp:199-204 (barycentric blend, flipped toward
    // the ray origin, normalised); sphere: (P - c)/r
    V3 n;
    bool backface = false;
    const Prim& pr = sc.prims[hit.prim];
    if (pr.is_sphere) {
      n = normalize(P - pr.v0);
    } else if (!sc.normals.empty()) {
      const float* nn = &sc.normals[(size_t)hit.prim * 9];
      float w0 = 1.0f - hit.u - hit.v;
      n = mk(nn[3], nn[4], nn[5]) * hit.u + mk(nn[6], nn[7], nn[8]) * hit.v + mk(nn[0], nn[1], nn[2]) * w0;
    } else {
      n = cross(pr.e1, pr.e2);
    }
    if (!(dot(d, n) < 0.0f)) { n = neg(n); backface = true; }
    if (!pr.is_sphere) n = normalize(n);
    V3 X, Y, Z;
    make_coord_space(n, &X, &Y, &Z);
    V3 wo_w = neg(d);
    V3 wo = mk(dot(wo_w, X), dot(wo_w, Y), dot(wo_w, Z));

    if (m.kind == B2RT_MAT_DIFFUSE || m.kind == B2RT_MAT_GLOSSY) {
      uint32_t j = 0;
      const size_t n_lights = sc.lights.size() + (sc.env_w ? 1 : 0);   // the environment map is one more light
      for (size_t li = 0; li < n_lights; ++li) {
        const bool is_env = li == sc.lights.size();
        static const b2rt_light no_light = {};
        const b2rt_light& lt = is_env ? no_light : sc.lights[li];
        uint32_t ns = (!is_env && lt.kind == B2RT_LIGHT_AREA) ? std::max(1u, rc.cfg.ns_area_light) : 1u;
        for (uint32_t k = 0; k < ns; ++k, ++j) {
          V3 wi; float dist, pdf; V3 Lr;
          V3 lp = mk(lt.position[0], lt.position[1], lt.position[2]);
          V3 ld = mk(lt.direction[0], lt.direction[1], lt.direction[2]);
          V3 rad = mk(lt.radiance[0], lt.radiance[1], lt.radiance[2]);
          if (is_env) {
            // EnvironmentLight::sample_L, uniform over the sphere: pdf = 1 / (4 pi)
            philox4x32_10(pix, sample, b, 1 + j, k0, k1, r4);
            const float zz = 1.0f - 2.0f * u01(r4[0]);
            float sn, cs;
            sincos2pi(u01(r4[1]), &sn, &cs);
            const float rr2 = 1.0f - zz * zz;
            const float rr = sqrtf(rr2 < 0.0f ? 0.0f : rr2);
            wi = mk(rr * cs, zz, rr * sn);
            dist = INF; pdf = 0.0795774715459476679f;
            Lr = env_lookup(sc, wi);
          } else if (lt.kind == B2RT_LIGHT_AREA) {
            philox4x32_10(pix, sample, b, 1 + j, k0, k1, r4);
            float ux = u01(r4[0]) - 0.5f, uy = u01(r4[1]) - 0.5f;
            V3 dv = lp + mk(lt.dim_x[0], lt.dim_x[1], lt.dim_x[2]) * ux + mk(lt.dim_y[0], lt.dim_y[1], lt.dim_y[2]) * uy - P;
            float sq = dot(dv, dv);
            dist = sqrtf(sq);
            float invd = 1.0f / dist;
            wi = dv * invd;
            float cosT = dot(wi, ld);
            pdf = sq / (rc.light_area[li] * fabsf(cosT));
            Lr = cosT < 0.0f ? rad : mk(0, 0, 0);
          } else if (lt.kind == B2RT_LIGHT_POINT) {
            V3 dv = lp - P;
            float sq = dot(dv, dv);
            dist = sqrtf(sq);
            wi = dv * (1.0f / dist);
            pdf = 1.0f; Lr = rad;
          } else {
            wi = neg(ld); dist = INF; pdf = 1.0f; Lr = rad;
          }
          float cos_in = dot(wi, Z);
          if (!(cos_in >= 0.0f)) continue;                       // pathtracer.cpp:462
          if (!(Lr.x > 0.0f || Lr.y > 0.0f || Lr.z > 0.0f)) continue;
          if (!(pdf > 0.0f)) continue;
          cn->rays_shadow++;
          if (occluded_bvh(sc, P, wi, rc.eps, dist - rc.eps, cn)) continue;
          float wgt = cos_in / ((float)ns * pdf);                  // pathtracer.cpp:473
          V3 f = bsdf_f(m, wo, mk(dot(wi, X), dot(wi, Y), cos_in));
          L = L + thr * f * Lr * wgt;
        }
      }
    }
    if (b + 1 == max_depth) break;
    // BSDF sample (the reference's sample_f bodies are stubs, src/bsdf.cpp:41-96; contracts in
    // src/bsdf.h): diffuse = cosine-weighted hemisphere, mirror = reflect about (0,0,1),
    // glass = Fresnel-weighted choice of reflect / refract, refraction = refract (TIR reflects).
    philox4x32_10(pix, sample, b, 0, k0, k1, r4);
    float u2 = u01(r4[2]), u3 = u01(r4[3]);
    V3 wi_l, weight;
    bool delta = false;
    if (m.kind == B2RT_MAT_DIFFUSE || m.kind == B2RT_MAT_GLOSSY) {
      float r = sqrtf(u2), s, c;
      sincos2pi(u3, &s, &c);
      float zz = 1.0f - u2;
      wi_l = mk(r * c, r * s, sqrtf(zz < 0.0f ? 0.0f : zz));
      // cosine-weighted sample, pdf = cos / pi: f * cos / pdf = f * pi (= albedo for the diffuse BSDF)
      weight = m.kind == B2RT_MAT_DIFFUSE ? mk(m.albedo[0], m.albedo[1], m.albedo[2]) : bsdf_f(m, wo, wi_l) * 3.14159265358979324f;
    } else if (m.kind == B2RT_MAT_MIRROR) {
      wi_l = mk(-wo.x, -wo.y, wo.z);
      weight = mk(m.albedo[0], m.albedo[1], m.albedo[2]);
      delta = true;
    } else {
      delta = true;
      float eta = backface ? m.ior : 1.0f / m.ior;   // n_incident / n_transmitted
      float cos_i = wo.z;
      float sin2_t = eta * eta * (1.0f - cos_i * cos_i);
      bool tir = !(sin2_t < 1.0f);
      float cos_t = tir ? 0.0f : sqrtf(1.0f - sin2_t);
      float Fr = 1.0f;
      if (!tir) {
        float ni = backface ? m.ior : 1.0f, nt = backface ? 1.0f : m.ior;
        float rs = (ni * cos_i - nt * cos_t) / (ni * cos_i + nt * cos_t);
        float rp = (nt * cos_i - ni * cos_t) / (nt * cos_i + ni * cos_t);
        Fr = 0.5f * (rs * rs + rp * rp);
      }
      bool reflect;
      if (m.kind == B2RT_MAT_GLASS) reflect = tir || (u2 < Fr);
      else reflect = tir;
      if (reflect) {
        wi_l = mk(-wo.x, -wo.y, wo.z);
        weight = m.kind == B2RT_MAT_GLASS ? mk(m.albedo[0], m.albedo[1], m.albedo[2]) : mk(1, 1, 1);
      } else {
        wi_l = mk(-wo.x * eta, -wo.y * eta, -cos_t);
        weight = mk(m.transmittance[0], m.transmittance[1], m.transmittance[2]);
      }
    }
    if (!(weight.x > 0.0f || weight.y > 0.0f || weight.z > 0.0f)) break;
    thr = thr * weight;
    d = normalize(X * wi_l.x + Y * wi_l.y + Z * wi_l.z);
    o = P;
    tmin = rc.eps;
    count_emission = delta;
  }
  return L;
}

Scene* make_scene(const b2rt_scene_desc* d) {
  Scene* sc = new Scene();
  sc->n_tris = d->n_tris; sc->n_spheres = d->n_spheres;
  sc->prims.resize((size_t)d->n_tris + d->n_spheres);
  sc->pbox.resize(sc->prims.size());
  sc->prim_mat.resize(sc->prims.size(), 0);
  const double PAD = 1e-3;  // triangle.cpp:36-44
  for (uint32_t i = 0; i < d->n_tris; ++i) {
    const float* v = d->tri_verts + (size_t)i * 9;
    Prim p;
    p.v0 = mk(v[0], v[1], v[2]);
    p.e1 = mk(v[3], v[4], v[5]) - p.v0;
    p.e2 = mk(v[6], v[7], v[8]) - p.v0;
    p.id = i; p.is_sphere = 0;
    sc->prims[i] = p;
    BBoxD b;
    for (int a = 0; a < 3; ++a) {
      double lo = std::min({(double)v[a], (double)v[3 + a], (double)v[6 + a]});
      double hi = std::max({(double)v[a], (double)v[3 + a], (double)v[6 + a]});
      b.mn[a] = lo - PAD; b.mx[a] = hi + PAD;
    }
    sc->pbox[i] = b;
    if (d->tri_material) sc->prim_mat[i] = d->tri_material[i];
  }
  for (uint32_t i = 0; i < d->n_spheres; ++i) {
    const float* s = d->spheres + (size_t)i * 4;
    Prim p;
    p.v0 = mk(s[0], s[1], s[2]); p.e1 = mk(s[3], 0, 0); p.e2 = mk(0, 0, 0);
    p.id = d->n_tris + i; p.is_sphere = 1;
    sc->prims[p.id] = p;
    BBoxD b;
    for (int a = 0; a < 3; ++a) { b.mn[a] = (double)s[a] - (double)s[3]; b.mx[a] = (double)s[a] + (double)s[3]; }
    sc->pbox[p.id] = b;
    if (d->sphere_material) sc->prim_mat[p.id] = d->sphere_material[i];
  }
  if (d->tri_normals) sc->normals.assign(d->tri_normals, d->tri_normals + (size_t)d->n_tris * 9);
  sc->mats.assign(d->materials, d->materials + d->n_materials);
  if (sc->mats.empty()) {
    b2rt_material m; memset(&m, 0, sizeof m); m.kind = B2RT_MAT_DIFFUSE; m.albedo[0] = m.albedo[1] = m.albedo[2] = 0.5f; m.ior = 1;
    sc->mats.push_back(m);
  }
  if (d->lights) sc->lights.assign(d->lights, d->lights + d->n_lights);
  return sc;
}

}  // namespace

extern "C" {

void* orc_scene_create(const b2rt_scene_desc* d, uint32_t max_leaf) {
  Scene* sc = make_scene(d);
  if (max_leaf != 0xFFFFFFFFu) build_bvh(*sc, max_leaf ? max_leaf : 4);   // 0xFFFFFFFF: exhaustive queries only (10 M soup)
  return sc;
}
void orc_scene_destroy(void* s) { delete (Scene*)s; }
// environment map for orc_render (rgb = nullptr removes it)
void orc_set_envmap(void* s, const float* rgb, uint32_t w, uint32_t h) {
  Scene* sc = (Scene*)s;
  sc->env.clear(); sc->env_w = sc->env_h = 0;
  if (rgb && w && h) { sc->env.assign(rgb, rgb + (size_t)w * h * 3); sc->env_w = w; sc->env_h = h; }
}
float orc_atan2_turns(float y, float x) { return atan2_turns(y, x); }

// BVH structure dump for the cross-check against oracle/_ref (the reference's own builder).
uint32_t orc_bvh_node_count(void* s) { return (uint32_t)((Scene*)s)->nodes.size(); }
void orc_bvh_dump(void* s, uint32_t* order, uint64_t* start, uint64_t* range, int32_t* left, int32_t* right) {
  Scene* sc = (Scene*)s;
  for (size_t i = 0; i < sc->order.size(); ++i) order[i] = sc->order[i];
  for (size_t i = 0; i < sc->nodes.size(); ++i) {
    start[i] = sc->nodes[i].start; range[i] = sc->nodes[i].range; left[i] = sc->nodes[i].l; right[i] = sc->nodes[i].r;
  }
}
uint32_t orc_wide_levels(void* s, uint32_t* levels, uint32_t cap) {
  Scene* sc = (Scene*)s;
  std::vector<uint32_t> lv;
  if (!sc->nodes.empty()) wide_levels(*sc, 0, 0, lv);
  for (size_t i = 0; i < lv.size() && i < cap; ++i) levels[i] = lv[i];
  return (uint32_t)lv.size();
}

// mode 0: exhaustive over all primitives; mode 1: BVH traversal.  any_hit -> hit_prim = 0/1 flag
void orc_intersect(void* s, int mode, int any_hit, const float* org, const float* dir, const float* tmin,
                   const float* tmax, uint64_t n, float* hit_t, uint32_t* hit_prim, int threads,
                   uint64_t* counters /* [box_tests, prim_tests] or NULL */) {
  Scene* sc = (Scene*)s;
  if (threads < 1) threads = 1;
  std::vector<Counters> cns(threads);
  std::atomic<uint64_t> next(0);
  auto work = [&](int ti) {
    const uint64_t chunk = 256;
    for (;;) {
      uint64_t b = next.fetch_add(chunk);
      if (b >= n) break;
      uint64_t e = std::min(n, b + chunk);
      for (uint64_t i = b; i < e; ++i) {
        V3 o = mk(org[3 * i], org[3 * i + 1], org[3 * i + 2]), d = mk(dir[3 * i], dir[3 * i + 1], dir[3 * i + 2]);
        Hit h;
        if (mode == 0) closest_brute(*sc, o, d, tmin[i], tmax[i], &h, &cns[ti]);
        else closest_bvh(*sc, o, d, tmin[i], tmax[i], &h, &cns[ti]);
        if (any_hit) { hit_prim[i] = h.prim != 0xFFFFFFFFu; if (hit_t) hit_t[i] = h.t; }
        else { hit_t[i] = h.t; hit_prim[i] = h.prim; }
      }
    }
  };
  std::vector<std::thread> th;
  for (int t = 1; t < threads; ++t) th.emplace_back(work, t);
  work(0);
  for (auto& t : th) t.join();
  if (counters) {
    Counters tot; for (auto& c : cns) tot.add(c);
    counters[0] = tot.box_tests; counters[1] = tot.prim_tests;
  }
}

// Render: PathTracer::start_raytracing + worker_thread + raytrace_tile (src/pathtracer.cpp:183-213,
// 510-558): 32x32 tiles (imageTileSize, :55) from a shared queue drained by num_threads threads.
// rgb: HDR buffer, Spectrum per pixel, index x + y*w (src/image.h:114-118), mean over ns_aa samples.
// out_stats: [rays_camera, rays_bounce, rays_shadow, box_tests, prim_tests], seconds.
// pixel_stride/pixel_offset restrict the run to a bounded sample of tiles (for the timed baseline).
int orc_render(void* s, const b2rt_camera* cam, const b2rt_config* cfg, uint32_t w, uint32_t h, int threads,
               float* rgb, uint64_t* out_stats, double* seconds, uint32_t tile_stride) {
  Scene* sc = (Scene*)s;
  RenderCtx rc;
  rc.sc = sc; rc.cfg = *cfg; rc.w = w; rc.h = h;
  rc.eps = cfg->ray_eps > 0 ? cfg->ray_eps : 1e-4f;
  rc.cam.pos = mk(cam->pos[0], cam->pos[1], cam->pos[2]);
  rc.cam.cx = mk(cam->c2w[0], cam->c2w[1], cam->c2w[2]);
  rc.cam.cy = mk(cam->c2w[3], cam->c2w[4], cam->c2w[5]);
  rc.cam.cz = mk(cam->c2w[6], cam->c2w[7], cam->c2w[8]);
  rc.cam.tan_h = tanf(cam->hfov_deg * 0.5f * 0.01745329251994329577f);
  rc.cam.tan_v = tanf(cam->vfov_deg * 0.5f * 0.01745329251994329577f);
  const uint32_t stride = cfg->sample_stride ? cfg->sample_stride : 1;
  rc.spp_total_for_jitter = cfg->ns_aa * stride;
  for (auto& l : sc->lights) {
    V3 dx = mk(l.dim_x[0], l.dim_x[1], l.dim_x[2]), dy = mk(l.dim_y[0], l.dim_y[1], l.dim_y[2]);
    rc.light_area.push_back(sqrtf(dot(dx, dx)) * sqrtf(dot(dy, dy)));
  }
  if (threads < 1) threads = 1;
  if (tile_stride < 1) tile_stride = 1;
  const uint32_t T = 32;
  const uint32_t tw = (w + T - 1) / T, th_ = (h + T - 1) / T;
  std::atomic<uint32_t> next(0);
  std::vector<Counters> cns(threads);
  auto t0 = std::chrono::steady_clock::now();
  auto work = [&](int ti) {
    for (;;) {
      uint32_t tile = next.fetch_add(1) * tile_stride;
      if (tile >= tw * th_) break;
      uint32_t x0 = (tile % tw) * T, y0 = (tile / tw) * T;
      for (uint32_t y = y0; y < std::min(h, y0 + T); ++y)
        for (uint32_t x = x0; x < std::min(w, x0 + T); ++x) {
          uint32_t pix = x + y * w;
          V3 sum = mk(0, 0, 0);
          for (uint32_t k = 0; k < cfg->ns_aa; ++k) {
            V3 L = trace_path(rc, pix, x, y, cfg->sample_first + k * stride, &cns[ti]);
            sum = sum + L;
          }
          float inv = 1.0f / (float)cfg->ns_aa;
          rgb[3 * (size_t)pix] = sum.x * inv; rgb[3 * (size_t)pix + 1] = sum.y * inv; rgb[3 * (size_t)pix + 2] = sum.z * inv;
        }
    }
  };
  std::vector<std::thread> ths;
  for (int t = 1; t < threads; ++t) ths.emplace_back(work, t);
  work(0);
  for (auto& t : ths) t.join();
  auto t1 = std::chrono::steady_clock::now();
  if (seconds) *seconds = std::chrono::duration<double>(t1 - t0).count();
  if (out_stats) {
    Counters tot; for (auto& c : cns) tot.add(c);
    out_stats[0] = tot.rays_camera; out_stats[1] = tot.rays_bounce; out_stats[2] = tot.rays_shadow;
    out_stats[3] = tot.box_tests; out_stats[4] = tot.prim_tests;
  }
  return 0;
}

// toColor + update_pixel, src/image.h:49-58,173-188
void orc_tonemap(const float* rgb, uint32_t n_pixels, uint32_t* rgba8) {
  float gamma = 2.2f, level = 1.0f, one_over_gamma = 1.0f / gamma;
  float exposure = sqrtf(powf(2, level));
  auto q = [](float c) { c = c < 0.f ? 0.f : (c > 1.f ? 1.f : c); return (uint32_t)(c * 255); };
  for (uint32_t i = 0; i < n_pixels; ++i) {
    float r = powf(rgb[3 * i] * exposure, one_over_gamma), g = powf(rgb[3 * i + 1] * exposure, one_over_gamma),
          b = powf(rgb[3 * i + 2] * exposure, one_over_gamma);
    rgba8[i] = (255u << 24) | (q(b) << 16) | (q(g) << 8) | q(r);
  }
}

// 3x3 per-channel median, border = 1.0, restating kernelMedianFilter (src/cudaRenderer.cu:773-842)
// on the row-major x + y*w layout.
void orc_median3x3(const float* rgb, uint32_t w, uint32_t h, float* out) {
  for (uint32_t y = 0; y < h; ++y)
    for (uint32_t x = 0; x < w; ++x)
      for (int c = 0; c < 3; ++c) {
        float v[9]; int k = 0;
        for (int dy = -1; dy <= 1; ++dy)
          for (int dx = -1; dx <= 1; ++dx) {
            int xx = (int)x + dx, yy = (int)y + dy;
            v[k++] = (xx < 0 || yy < 0 || xx >= (int)w || yy >= (int)h) ? 1.0f : rgb[3 * ((size_t)xx + (size_t)yy * w) + c];
          }
        std::sort(v, v + 9);
        out[3 * ((size_t)x + (size_t)y * w) + c] = v[4];
      }
}

// Gaussian (kind 1: 3x3 binomial) / joint bilateral (kind 2: 5x5 binomial x 1/(1 + |dc|^2/sigma_r^2)) reconstruction
// filters, b2rt_config.filter_kind.  No reference implementation exists (the reference's Gaussian is commented out,
// src/cudaRenderer.cu:755-771); this restatement fixes the definition: taps outside the image are dropped, weights
// renormalised, taps accumulated in row-major order with separate multiply and add.
void orc_filter(uint32_t kind, float sigma_r, const float* rgb, uint32_t w, uint32_t h, float* out) {
  const int R = kind == 2 ? 2 : 1;
  const float sr = sigma_r > 0.f ? sigma_r : 0.25f;
  const float inv_s2 = 1.0f / (sr * sr);
  static const int B1[3] = {1, 2, 1}, B2[5] = {1, 4, 6, 4, 1};
  for (uint32_t y = 0; y < h; ++y)
    for (uint32_t x = 0; x < w; ++x) {
      const float* c = rgb + 3 * ((size_t)x + (size_t)y * w);
      float a[3] = {0.f, 0.f, 0.f}, ws = 0.f;
      for (int dy = -R; dy <= R; ++dy)
        for (int dx = -R; dx <= R; ++dx) {
          const int xx = (int)x + dx, yy = (int)y + dy;
          if (xx < 0 || yy < 0 || xx >= (int)w || yy >= (int)h) continue;
          const float* q = rgb + 3 * ((size_t)xx + (size_t)yy * w);
          const float bw = R == 1 ? (float)(B1[dx + 1] * B1[dy + 1]) : (float)(B2[dx + 2] * B2[dy + 2]);
          float wgt = bw;
          if (kind == 2) {
            const float dr = q[0] - c[0], dg = q[1] - c[1], db = q[2] - c[2];
            const float d2 = (dr * dr + dg * dg) + db * db;
            wgt = bw * (1.0f / (1.0f + d2 * inv_s2));
          }
          for (int k = 0; k < 3; ++k) a[k] = a[k] + wgt * q[k];
          ws = ws + wgt;
        }
      for (int k = 0; k < 3; ++k) out[3 * ((size_t)x + (size_t)y * w) + k] = a[k] / ws;
    }
}

// known-answer helpers for tests
void orc_philox(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t* out4) {
  philox4x32_10(c0, c1, c2, c3, k0, k1, out4);
}
void orc_sincos2pi(float u, float* s, float* c) { sincos2pi(u, s, c); }

}  // extern "C"
