// Host BVH construction for the B200 traversal path.
//
// Replaces, with a different algorithm and output layout, the reference's
//   BVHAccel::BVHAccel + splitBVHNode      src/bvh.cpp:339-365, 48-230   (binary SAH build)
//   BVHNode::compactTree                   src/bvh.cpp:275-337           (collapse to 4-wide)
//   BVHSubTree::compress                   src/bvh.cpp:234-273           (serialise + level lists)
//   CuTriangle / CuBVHSubTree upload       src/cudaRenderer.cu:1757-1827
// Pipeline here: binned-SAH binary build (O(n log n), threaded) -> SAH-ordered collapse to a W-wide
// tree (W = 4 or 8) -> partition into shared-memory-sized subtrees ("treelets") in BFS order ->
// one position-independent blob per subtree: SoA wide nodes (128-bit aligned rows) followed by the
// 48-byte primitive records of its leaves.  Closest-hit results do not depend on the tree shape, so
// this builder does not have to reproduce the reference's tree (the oracle's builder does).
#include <emmintrin.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <queue>
#include <thread>

#include "b2rt_internal.h"

namespace b2rt {

namespace {

struct Box {
  float mn[3], mx[3];
  void reset() {
    for (int a = 0; a < 3; ++a) { mn[a] = std::numeric_limits<float>::infinity(); mx[a] = -std::numeric_limits<float>::infinity(); }
  }
  void grow(const Box& b) {
    for (int a = 0; a < 3; ++a) { mn[a] = std::min(mn[a], b.mn[a]); mx[a] = std::max(mx[a], b.mx[a]); }
  }
  void grow(const float* p) {
    for (int a = 0; a < 3; ++a) { mn[a] = std::min(mn[a], p[a]); mx[a] = std::max(mx[a], p[a]); }
  }
  float area() const {
    float ex = mx[0] - mn[0], ey = mx[1] - mn[1], ez = mx[2] - mn[2];
    if (ex < 0 || ey < 0 || ez < 0) return 0.f;
    return 2.f * (ex * ey + ey * ez + ez * ex);
  }
};

struct BinNode {
  Box box;
  uint32_t left = 0, right = 0;   // children (0 = none => leaf)
  uint32_t start = 0, count = 0;  // primitive range in the index array
};

// ---- worker pool of the host builder ---------------------------------------------------------------------------
// One lazily created pool per process.  Workers sleep on a condition variable between builds and spin while a build
// is running (a build is a few milliseconds of many short parallel regions; a condition-variable wake-up per region
// would cost more than the region).  parallel_for hands out items one at a time under the pool mutex -- items are
// coarse (thousands of primitives each).
class BuildPool {
 public:
  static BuildPool& get() { static BuildPool p; return p; }
  int workers() const { return (int)threads_.size(); }
  // a build owns the pool from begin() to end(); a second concurrent build in the same process runs without it
  bool begin() {
    if (threads_.empty() || !owner_.try_lock()) return false;
    { std::lock_guard<std::mutex> lk(m_); hot_ = true; hot_flag_.store(true); }
    cv_.notify_all();
    return true;
  }
  void end() {
    { std::lock_guard<std::mutex> lk(m_); hot_ = false; hot_flag_.store(false); }
    owner_.unlock();
  }
  void parallel_for(uint32_t n, const std::function<void(uint32_t)>& fn) {
    if (n == 0) return;
    {
      std::lock_guard<std::mutex> lk(m_);
      job_ = &fn; n_ = n; next_ = 0; remaining_.store(n); epoch_.fetch_add(1, std::memory_order_release);
    }
    for (;;) {   // the caller works too
      uint32_t i;
      { std::lock_guard<std::mutex> lk(m_); if (next_ >= n_) break; i = next_++; }
      fn(i);
      remaining_.fetch_sub(1);
    }
    while (remaining_.load() != 0) _mm_pause();
  }

 private:
  BuildPool() {
    unsigned hw = std::max(1u, std::thread::hardware_concurrency());
    unsigned share = 1;   // ranks of one node share the host cores (torchrun exports LOCAL_WORLD_SIZE)
    if (const char* e = getenv("LOCAL_WORLD_SIZE")) share = (unsigned)std::max(1, atoi(e));
    unsigned want = std::min(16u, std::max(1u, hw / share));
    if (const char* e = getenv("B2RT_BUILD_THREADS")) want = (unsigned)std::max(1, atoi(e));
    for (unsigned t = 1; t < want; ++t) threads_.emplace_back([this] { worker(); });
  }
  ~BuildPool() {
    { std::lock_guard<std::mutex> lk(m_); stop_ = true; }
    cv_.notify_all();
    for (auto& t : threads_) t.join();
  }
  void worker() {
    uint64_t seen = 0;
    std::unique_lock<std::mutex> lk(m_);
    for (;;) {
      while (epoch_.load() == seen && !stop_) {
        if (hot_) {   // a build is running: spin on the epoch without touching the mutex
          lk.unlock();
          while (epoch_.load(std::memory_order_acquire) == seen && hot_flag_.load(std::memory_order_relaxed)) _mm_pause();
          lk.lock();
        } else cv_.wait(lk);
      }
      if (stop_) return;
      seen = epoch_.load();
      while (epoch_.load() == seen && next_ < n_) {
        const uint32_t i = next_++;
        const std::function<void(uint32_t)>* job = job_;
        lk.unlock();
        (*job)(i);
        remaining_.fetch_sub(1);
        lk.lock();
      }
    }
  }
  std::vector<std::thread> threads_;
  std::mutex m_, owner_;
  std::condition_variable cv_;
  const std::function<void(uint32_t)>* job_ = nullptr;
  uint32_t n_ = 0, next_ = 0;
  std::atomic<uint64_t> epoch_{0};
  std::atomic<uint32_t> remaining_{0};
  std::atomic<bool> hot_flag_{false};
  bool hot_ = false, stop_ = false;
};

// Binned-SAH binary builder.  The per-primitive passes (bounds, 3-axis binning, partition) run on SSE registers: a box
// is two __m128 (min xyz_, max xyz_), the centroid is recomputed instead of stored, and all three axes are binned in
// the same pass over the primitives.  Large nodes run their passes in parallel over chunks of the index range (the
// merged bounds / bins / stable partition are exactly what the serial passes produce, so the tree does not depend on
// the thread count); the subtrees below them are then built independently, one pool item each.
struct BinaryBuilder {
  static constexpr int NBMAX = 16;
  struct Bounds { __m128 nmn, nmx, cmn, cmx; };
  struct Bins { __m128 mn[3][NBMAX], mx[3][NBMAX]; uint32_t c[3][NBMAX]; };
  struct Split { int axis = -1, bin = -1, nb = 0; float lo = 0.f, sc = 0.f; };

  std::vector<__m128> pmn, pmx;    // primitive boxes, lane 3 unused
  std::vector<uint32_t> idx, scratch_l, scratch_r;
  std::vector<BinNode> nodes;
  std::atomic<uint32_t> next_node{0};
  uint32_t max_leaf;
  float sah_ct = 1.5f;    // cost of one more child box in units of a primitive test (measured, profiles/r02_sweep_sah_ct.txt; 1e30 = never split a range that fits a leaf)

  BinaryBuilder(const std::vector<Box>& pb, uint32_t ml) : max_leaf(ml) {
    if (const char* e = getenv("B2RT_SAH_CT")) sah_ct = (float)atof(e);
    size_t n = pb.size();
    pmn.resize(n); pmx.resize(n);
    idx.resize(n); scratch_l.resize(n + 1); scratch_r.resize(n + 1);
    for (size_t i = 0; i < n; ++i) {
      idx[i] = (uint32_t)i;
      pmn[i] = _mm_set_ps(0.f, pb[i].mn[2], pb[i].mn[1], pb[i].mn[0]);
      pmx[i] = _mm_set_ps(0.f, pb[i].mx[2], pb[i].mx[1], pb[i].mx[0]);
    }
    nodes.resize(std::max<size_t>(1, 2 * n));
  }

  static float area_of(__m128 mn, __m128 mx) {
    alignas(16) float e[4];
    _mm_store_ps(e, _mm_sub_ps(mx, mn));
    if (e[0] < 0 || e[1] < 0 || e[2] < 0) return 0.f;
    return 2.f * (e[0] * e[1] + e[1] * e[2] + e[2] * e[0]);
  }
  static Bounds empty_bounds() {
    const __m128 pinf = _mm_set1_ps(std::numeric_limits<float>::infinity()), ninf = _mm_set1_ps(-std::numeric_limits<float>::infinity());
    return Bounds{pinf, ninf, pinf, ninf};
  }
  static void merge(Bounds& a, const Bounds& b) {
    a.nmn = _mm_min_ps(a.nmn, b.nmn); a.nmx = _mm_max_ps(a.nmx, b.nmx); a.cmn = _mm_min_ps(a.cmn, b.cmn); a.cmx = _mm_max_ps(a.cmx, b.cmx);
  }
  Bounds bounds_range(uint32_t b, uint32_t e) const {
    Bounds r = empty_bounds();
    const __m128 half = _mm_set1_ps(0.5f);
    for (uint32_t i = b; i < e; ++i) {
      const uint32_t p = idx[i];
      const __m128 mn = pmn[p], mx = pmx[p], c = _mm_mul_ps(half, _mm_add_ps(mn, mx));
      r.nmn = _mm_min_ps(r.nmn, mn); r.nmx = _mm_max_ps(r.nmx, mx);
      r.cmn = _mm_min_ps(r.cmn, c); r.cmx = _mm_max_ps(r.cmx, c);
    }
    return r;
  }
  // bin geometry of a node: 16 bins for big nodes, fewer for small ones (most nodes are small; resetting and sweeping
  // 3 x 16 bins would dominate the build)
  struct BinSetup { int nb; bool use[3]; alignas(16) float lo3[4]; alignas(16) float scale3[4]; };
  static BinSetup bin_setup(const Bounds& bd, uint32_t count) {
    BinSetup s;
    s.nb = count >= 64 ? 16 : (count >= 16 ? 8 : 4);
    alignas(16) float hi3[4];
    _mm_store_ps(s.lo3, bd.cmn); _mm_store_ps(hi3, bd.cmx);
    for (int a = 0; a < 3; ++a) {
      const float ext = hi3[a] - s.lo3[a];
      s.use[a] = ext > 0.f;
      s.scale3[a] = s.use[a] ? (float)s.nb / ext : 0.f;
    }
    s.scale3[3] = 0.f;
    return s;
  }
  static void reset(Bins& B, int nb) {
    const __m128 pinf = _mm_set1_ps(std::numeric_limits<float>::infinity()), ninf = _mm_set1_ps(-std::numeric_limits<float>::infinity());
    for (int a = 0; a < 3; ++a)
      for (int k = 0; k < nb; ++k) { B.mn[a][k] = pinf; B.mx[a][k] = ninf; B.c[a][k] = 0; }
  }
  static void merge(Bins& A, const Bins& B, int nb) {
    for (int a = 0; a < 3; ++a)
      for (int k = 0; k < nb; ++k) { A.mn[a][k] = _mm_min_ps(A.mn[a][k], B.mn[a][k]); A.mx[a][k] = _mm_max_ps(A.mx[a][k], B.mx[a][k]); A.c[a][k] += B.c[a][k]; }
  }
  void bin_range(uint32_t b, uint32_t e, const Bounds& bd, const BinSetup& s, Bins& B) const {
    const __m128 half = _mm_set1_ps(0.5f), scale = _mm_load_ps(s.scale3), kmax = _mm_set1_ps((float)(s.nb - 1)), zero = _mm_setzero_ps();
    for (uint32_t i = b; i < e; ++i) {
      const uint32_t p = idx[i];
      const __m128 mn = pmn[p], mx = pmx[p], c = _mm_mul_ps(half, _mm_add_ps(mn, mx));
      const __m128 kf = _mm_min_ps(_mm_max_ps(_mm_mul_ps(_mm_sub_ps(c, bd.cmn), scale), zero), kmax);
      alignas(16) int k4[4];
      _mm_store_si128(reinterpret_cast<__m128i*>(k4), _mm_cvttps_epi32(kf));
      for (int a = 0; a < 3; ++a) {
        if (!s.use[a]) continue;
        const int k = k4[a];
        B.mn[a][k] = _mm_min_ps(B.mn[a][k], mn); B.mx[a][k] = _mm_max_ps(B.mx[a][k], mx); B.c[a][k]++;
      }
    }
  }
  static Split pick_split(const Bins& B, const BinSetup& s, float* cost_out = nullptr) {
    const __m128 pinf = _mm_set1_ps(std::numeric_limits<float>::infinity()), ninf = _mm_set1_ps(-std::numeric_limits<float>::infinity());
    Split sp; sp.nb = s.nb;
    float best_cost = std::numeric_limits<float>::infinity();
    const int NB = s.nb;
    for (int a = 0; a < 3; ++a) {
      if (!s.use[a]) continue;
      float ra[NBMAX]; uint32_t rc[NBMAX];
      __m128 amn = pinf, amx = ninf; uint32_t c = 0;
      for (int k = NB - 1; k > 0; --k) { amn = _mm_min_ps(amn, B.mn[a][k]); amx = _mm_max_ps(amx, B.mx[a][k]); c += B.c[a][k]; ra[k] = area_of(amn, amx); rc[k] = c; }
      amn = pinf; amx = ninf; c = 0;
      for (int k = 0; k < NB - 1; ++k) {
        amn = _mm_min_ps(amn, B.mn[a][k]); amx = _mm_max_ps(amx, B.mx[a][k]); c += B.c[a][k];
        if (c == 0 || rc[k + 1] == 0) continue;
        float cost = area_of(amn, amx) * (float)c + ra[k + 1] * (float)rc[k + 1];
        if (cost < best_cost) { best_cost = cost; sp.axis = a; sp.bin = k; }
      }
    }
    if (sp.axis >= 0) { sp.lo = s.lo3[sp.axis]; sp.sc = s.scale3[sp.axis]; }
    if (cost_out) *cost_out = best_cost;   // sum over the two sides of area x count
    return sp;
  }
  // branch-free stable partition of idx[b, e) into L / R scratch (the side of a primitive is a coin flip for the branch
  // predictor; std::partition spent a third of the build in mispredictions); returns the number that went left
  uint32_t split_range(uint32_t b, uint32_t e, const Split& sp, uint32_t* L, uint32_t* R) const {
    uint32_t nl = 0, nr = 0;
    const float kmaxf = (float)(sp.nb - 1);
    const int ax = sp.axis;
    for (uint32_t i = b; i < e; ++i) {
      const uint32_t p = idx[i];
      const float* mn = reinterpret_cast<const float*>(&pmn[p]); const float* mx = reinterpret_cast<const float*>(&pmx[p]);
      float kf = (0.5f * (mn[ax] + mx[ax]) - sp.lo) * sp.sc;
      kf = std::min(std::max(kf, 0.f), kmaxf);
      const uint32_t left = (int)kf <= sp.bin ? 1u : 0u;
      L[nl] = p; R[nr] = p;
      nl += left; nr += 1u - left;
    }
    return nl;
  }
  void set_node(uint32_t me, uint32_t b, uint32_t e, const Bounds& bd) {
    BinNode& nd = nodes[me];
    alignas(16) float t4[4];
    _mm_store_ps(t4, bd.nmn); nd.box.mn[0] = t4[0]; nd.box.mn[1] = t4[1]; nd.box.mn[2] = t4[2];
    _mm_store_ps(t4, bd.nmx); nd.box.mx[0] = t4[0]; nd.box.mx[1] = t4[1]; nd.box.mx[2] = t4[2];
    nd.start = b; nd.count = e - b; nd.left = nd.right = 0;
  }

  // serial recursive build of idx[b, e)
  uint32_t build(uint32_t b, uint32_t e) {
    const uint32_t me = next_node.fetch_add(1);
    const Bounds bd = bounds_range(b, e);
    set_node(me, b, e, bd);
    if (e - b <= 1) return me;
    const BinSetup bs = bin_setup(bd, e - b);
    Bins B; reset(B, bs.nb);
    bin_range(b, e, bd, bs, B);
    float split_cost = 0.f;
    const Split sp = pick_split(B, bs, &split_cost);
    if (e - b <= max_leaf) {
      // SAH leaf termination: a range that fits a leaf is still split when the expected primitive tests saved exceed
      // the cost of the extra box (sah_ct, in units of one primitive test).  What this catches: a few LARGE primitives
      // sharing a leaf (the walls of a box around a mesh) -- every ray that enters tests all of them.
      const float node_area = area_of(bd.nmn, bd.nmx);
      if (sp.axis < 0 || !(node_area > 0.f) || !(sah_ct + split_cost / node_area < (float)(e - b))) return me;
    }
    uint32_t mid = b;
    if (sp.axis >= 0) {
      uint32_t* L = scratch_l.data() + b; uint32_t* R = scratch_r.data() + b;
      const uint32_t nl = split_range(b, e, sp, L, R);
      memcpy(idx.data() + b, L, (size_t)nl * 4);
      memcpy(idx.data() + b + nl, R, (size_t)(e - b - nl) * 4);
      mid = b + nl;
    }
    if (mid == b || mid == e) mid = b + (e - b) / 2;   // all centroids coincide (or binning degenerate): split by index
    const uint32_t l = build(b, mid);
    const uint32_t r = build(mid, e);
    nodes[me].left = l; nodes[me].right = r;
    return me;
  }

  // one large node with its three passes spread over the pool; returns mid
  uint32_t split_node_parallel(BuildPool& pool, uint32_t me, uint32_t b, uint32_t e) {
    const uint32_t n = e - b;
    const uint32_t nchunk = std::min<uint32_t>((uint32_t)pool.workers() + 1, std::max(1u, n / 1024u));
    auto lo_of = [&](uint32_t c) { return b + (uint32_t)((uint64_t)n * c / nchunk); };
    std::vector<Bounds> cb(nchunk);
    pool.parallel_for(nchunk, [&](uint32_t c) { cb[c] = bounds_range(lo_of(c), lo_of(c + 1)); });
    Bounds bd = empty_bounds();
    for (auto& x : cb) merge(bd, x);
    set_node(me, b, e, bd);
    if (n <= max_leaf) return b;
    const BinSetup bs = bin_setup(bd, n);
    std::vector<Bins> bins(nchunk);
    pool.parallel_for(nchunk, [&](uint32_t c) { reset(bins[c], bs.nb); bin_range(lo_of(c), lo_of(c + 1), bd, bs, bins[c]); });
    for (uint32_t c = 1; c < nchunk; ++c) merge(bins[0], bins[c], bs.nb);
    const Split sp = pick_split(bins[0], bs);
    uint32_t mid = b;
    if (sp.axis >= 0) {
      // chunk c partitions into scratch[lo_of(c) ..); the left parts are then packed in chunk order, then the right parts
      std::vector<uint32_t> nl(nchunk);
      pool.parallel_for(nchunk, [&](uint32_t c) { nl[c] = split_range(lo_of(c), lo_of(c + 1), sp, scratch_l.data() + lo_of(c), scratch_r.data() + lo_of(c)); });
      std::vector<uint32_t> offl(nchunk), offr(nchunk);
      uint32_t tl = 0;
      for (uint32_t c = 0; c < nchunk; ++c) { offl[c] = tl; tl += nl[c]; }
      uint32_t tr = 0;
      for (uint32_t c = 0; c < nchunk; ++c) { offr[c] = tl + tr; tr += (lo_of(c + 1) - lo_of(c)) - nl[c]; }
      pool.parallel_for(nchunk, [&](uint32_t c) {
        memcpy(idx.data() + b + offl[c], scratch_l.data() + lo_of(c), (size_t)nl[c] * 4);
        memcpy(idx.data() + b + offr[c], scratch_r.data() + lo_of(c), (size_t)((lo_of(c + 1) - lo_of(c)) - nl[c]) * 4);
      });
      mid = b + tl;
    }
    if (mid == b || mid == e) mid = b + n / 2;
    return mid;
  }

  void build_all(uint32_t n) {
    BuildPool& pool = BuildPool::get();
    uint32_t serial_below = 4096;   // subtrees this small are pool items, built by build()
    if (const char* e = getenv("B2RT_BUILD_SERIAL_BELOW")) serial_below = (uint32_t)std::max(256, atoi(e));
    if (n <= serial_below || !pool.begin()) { build(0, n); return; }
    struct Item { uint32_t b, e, parent, side; };   // side 0 / 1 = left / right of `parent`; parent 0xFFFFFFFF = root
    std::vector<Item> level{{0u, n, 0xFFFFFFFFu, 0u}}, small;
    while (!level.empty()) {
      std::vector<Item> next;
      for (const Item& it : level) {
        if (it.e - it.b <= serial_below) { small.push_back(it); continue; }
        const uint32_t me = next_node.fetch_add(1);
        if (it.parent != 0xFFFFFFFFu) (it.side ? nodes[it.parent].right : nodes[it.parent].left) = me;
        const uint32_t mid = split_node_parallel(pool, me, it.b, it.e);
        next.push_back({it.b, mid, me, 0u});
        next.push_back({mid, it.e, me, 1u});
      }
      level.swap(next);
    }
    std::vector<uint32_t> res(small.size());
    pool.parallel_for((uint32_t)small.size(), [&](uint32_t k) { res[k] = build(small[k].b, small[k].e); });
    for (size_t k = 0; k < small.size(); ++k) (small[k].side ? nodes[small[k].parent].right : nodes[small[k].parent].left) = res[k];
    pool.end();
  }
};

struct WChild {
  Box box;
  int32_t node = -1;            // wide node index, or -1 for a leaf
  uint32_t start = 0, count = 0;  // leaf primitive range (index array)
};
struct WNode {
  std::vector<WChild> ch;
  float area = 0;
  uint64_t subtree_bytes = 0;   // this node + everything below
  uint32_t height = 1;          // levels of wide nodes below and including this one
  uint64_t own_bytes = 0;       // node record + its leaf primitives
  uint64_t subtree_nodes = 1;   // wide nodes below and including this one
};

}  // namespace

int make_host_scene(const b2rt_scene_desc* d, HostScene* out, bool with_geometry) {
  if (!d) { set_error("scene desc is null"); return B2RT_ERR_INVALID; }
  if (d->n_tris && !d->tri_verts) { set_error("tri_verts is null"); return B2RT_ERR_INVALID; }
  if (d->n_spheres && !d->spheres) { set_error("spheres is null"); return B2RT_ERR_INVALID; }
  if ((uint64_t)d->n_tris + d->n_spheres >= 0xFFFFFFF0ull) { set_error("too many primitives"); return B2RT_ERR_INVALID; }
  out->n_tris = d->n_tris; out->n_spheres = d->n_spheres;
  const uint32_t n = out->n_prims();
  out->prim_material.assign(n, 0u);
  if (d->tri_material) memcpy(out->prim_material.data(), d->tri_material, (size_t)d->n_tris * 4);
  if (d->sphere_material) memcpy(out->prim_material.data() + d->n_tris, d->sphere_material, (size_t)d->n_spheres * 4);
  out->prim_geom.clear();
  out->tri_normals.clear();
  if (with_geometry) {   // (the device builder makes the same records from the caller's arrays on the GPU: k_make_prims)
    out->prim_geom.assign((size_t)n * 12, 0.f);
    for (uint32_t i = 0; i < d->n_tris; ++i) {
      const float* v = d->tri_verts + (size_t)i * 9;
      float* g = &out->prim_geom[(size_t)i * 12];
      g[0] = v[0]; g[1] = v[1]; g[2] = v[2];
      g[3] = v[3] - v[0]; g[4] = v[4] - v[1]; g[5] = v[5] - v[2];   // e1 = p2 - p1 (triangle.cpp:172)
      g[6] = v[6] - v[0]; g[7] = v[7] - v[1]; g[8] = v[8] - v[2];   // e2 = p3 - p1
      uint32_t id = i, kind = 0;
      memcpy(&g[9], &id, 4); memcpy(&g[10], &kind, 4);
    }
    for (uint32_t i = 0; i < d->n_spheres; ++i) {
      const float* s = d->spheres + (size_t)i * 4;
      float* g = &out->prim_geom[((size_t)d->n_tris + i) * 12];
      g[0] = s[0]; g[1] = s[1]; g[2] = s[2]; g[3] = s[3];
      uint32_t id = d->n_tris + i, kind = 1;
      memcpy(&g[9], &id, 4); memcpy(&g[10], &kind, 4);
    }
    if (d->tri_normals) out->tri_normals.assign(d->tri_normals, d->tri_normals + (size_t)d->n_tris * 9);
  }
  if (d->n_materials && d->materials) out->materials.assign(d->materials, d->materials + d->n_materials);
  else {
    b2rt_material m; memset(&m, 0, sizeof m);
    m.kind = B2RT_MAT_DIFFUSE; m.albedo[0] = m.albedo[1] = m.albedo[2] = 0.5f; m.ior = 1.f;
    out->materials.assign(1, m);
  }
  if (out->materials.size() >= (1u << 28)) { set_error("too many materials"); return B2RT_ERR_INVALID; }
  for (const b2rt_material& m : out->materials)
    if (m.kind < B2RT_MAT_DIFFUSE || m.kind > B2RT_MAT_GLOSSY) { set_error("unknown material kind"); return B2RT_ERR_INVALID; }
  for (uint32_t i = 0; i < n; ++i)
    if (out->prim_material[i] >= out->materials.size()) { set_error("material index out of range"); return B2RT_ERR_INVALID; }
  if (d->n_lights && d->lights) out->lights.assign(d->lights, d->lights + d->n_lights);
  else out->lights.clear();
  return B2RT_OK;
}

int build_wide_bvh(const HostScene& sc, uint32_t max_leaf, uint32_t width, uint32_t treelet_bytes, WideBVH* out) {
  auto t0 = std::chrono::steady_clock::now();
  if (width == 0) width = 4;
  if (!width_ok(width)) { set_error("bvh width must be 2, 4, 8 or 16"); return B2RT_ERR_INVALID; }
  if (max_leaf == 0) max_leaf = 4;
  if (max_leaf > 64) { set_error("max_leaf_size must be <= 64"); return B2RT_ERR_INVALID; }
  const uint32_t NB = node_bytes(width);
  const uint32_t min_budget = NB + width * max_leaf * PRIM_BYTES;
  // default subtree budget, by measurement (round 2 kernel, profiles/r02_sweep_subtree_chunk.txt): 20 KiB for small
  // scenes (cfg2, 28 K triangles; 4 CTAs per SM), 24 KiB from 64 K primitives on (cfg3 stand-in, 114 K; 3 CTAs per SM)
  if (treelet_bytes == 0) treelet_bytes = (sc.n_prims() >= 65536 ? 24 : 20) * 1024;
  treelet_bytes = std::max(treelet_bytes, min_budget) & ~15u;
  // 227 KiB per CTA = subtree blob + per-thread stacks (<= 35 KiB) + the push staging rings
  if (treelet_bytes > 160 * 1024) { set_error("treelet_bytes exceeds the shared-memory budget (160 KiB)"); return B2RT_ERR_INVALID; }
  // a ray pushes at most width-1 children per wide node it descends through, so a subtree of at most depth_limit
  // node levels never needs more than stack_entries(width) stack slots (the kernel's shared-memory stack)
  const uint32_t depth_limit = stack_entries(width) / (width - 1);
  const uint64_t node_limit = max_treelet_nodes(width);

  const uint32_t n = sc.n_prims();
  out->width = width;
  out->blob.clear(); out->treelets.clear(); out->levels.clear();

  // primitive boxes, padded so the fp32 slab test stays conservative w.r.t. the primitive tests
  std::vector<Box> pbox(n);
  Box scene_box; scene_box.reset();
  double projected = 0;   // sum of the primitives' direction-averaged projected areas (triangle: area / 2)
  for (uint32_t i = 0; i < n; ++i) {
    const float* g = &sc.prim_geom[(size_t)i * 12];
    Box b; b.reset();
    if (i < sc.n_tris) {
      float p1[3] = {g[0], g[1], g[2]}, p2[3] = {g[0] + g[3], g[1] + g[4], g[2] + g[5]}, p3[3] = {g[0] + g[6], g[1] + g[7], g[2] + g[8]};
      b.grow(p1); b.grow(p2); b.grow(p3);
      const double cx = (double)g[4] * g[8] - (double)g[5] * g[7], cy = (double)g[5] * g[6] - (double)g[3] * g[8],
                   cz = (double)g[3] * g[7] - (double)g[4] * g[6];
      projected += 0.25 * std::sqrt(cx * cx + cy * cy + cz * cz);
    } else {
      for (int a = 0; a < 3; ++a) { b.mn[a] = g[a] - g[3]; b.mx[a] = g[a] + g[3]; }
      projected += 3.14159265358979 * (double)g[3] * g[3];
    }
    pbox[i] = b;
    scene_box.grow(b);
  }
  float diag = 0.f, maxabs = 0.f;
  if (n) {
    float ex = scene_box.mx[0] - scene_box.mn[0], ey = scene_box.mx[1] - scene_box.mn[1], ez = scene_box.mx[2] - scene_box.mn[2];
    diag = std::sqrt(ex * ex + ey * ey + ez * ez);
    for (int a = 0; a < 3; ++a) maxabs = std::max(maxabs, std::max(std::fabs(scene_box.mn[a]), std::fabs(scene_box.mx[a])));
  }
  const float pad = std::max(1e-30f, 1e-5f * std::max(diag, maxabs));
  for (auto& b : pbox)
    for (int a = 0; a < 3; ++a) { b.mn[a] -= pad; b.mx[a] += pad; }
  for (int a = 0; a < 3; ++a) { out->bbox[a] = n ? scene_box.mn[a] : 0.f; out->bbox[3 + a] = n ? scene_box.mx[a] : 0.f; }
  // mean free path of a random ray through the scene box if the primitives were spread evenly in it: volume / summed
  // projected area.  Only a scale for the distance slices of Tracer::trace_sliced, never a correctness input.
  out->mean_free_path = 0.f;
  if (n && projected > 0) {
    const double vol = (double)(scene_box.mx[0] - scene_box.mn[0]) * (scene_box.mx[1] - scene_box.mn[1]) * (scene_box.mx[2] - scene_box.mn[2]);
    out->mean_free_path = (float)(vol / projected);
  }

  const bool verbose = getenv("B2RT_VERBOSE") != nullptr;
  auto lap = [&](const char* what) {
    if (verbose) fprintf(stderr, "b2rt: build %-10s %8.3f ms\n", what, std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count());
  };
  lap("boxes");
  // ---- binary build ----
  BinaryBuilder bb(pbox, max_leaf);
  if (n) bb.build_all(n);
  lap("binary");

  // ---- collapse to W-wide ----
  std::vector<WNode> wn;
  if (n) {
    // iterative: (binary node) -> wide node
    struct Item { uint32_t bin; int32_t wide; };
    std::vector<Item> todo;
    wn.emplace_back();
    todo.push_back({0u, 0});
    while (!todo.empty()) {
      Item it = todo.back(); todo.pop_back();
      const BinNode& root = bb.nodes[it.bin];
      std::vector<uint32_t> kids;   // binary node ids forming the wide node's children
      if (root.left == 0 && root.right == 0) kids.push_back(it.bin);
      else { kids.push_back(root.left); kids.push_back(root.right); }
      while (kids.size() < width) {
        int best = -1; float ba = -1.f;
        for (size_t k = 0; k < kids.size(); ++k) {
          const BinNode& c = bb.nodes[kids[k]];
          if (c.left == 0 && c.right == 0) continue;
          float a = c.box.area();
          if (a > ba) { ba = a; best = (int)k; }
        }
        if (best < 0) break;
        const BinNode& c = bb.nodes[kids[best]];
        uint32_t l = c.left, r = c.right;
        kids[best] = l; kids.push_back(r);
      }
      std::vector<WChild> ch;
      for (uint32_t k : kids) {
        const BinNode& c = bb.nodes[k];
        WChild w; w.box = c.box;
        if (c.left == 0 && c.right == 0) { w.node = -1; w.start = c.start; w.count = c.count; }
        else {
          w.node = (int32_t)wn.size();
          wn.emplace_back();
          todo.push_back({k, w.node});
        }
        ch.push_back(w);
      }
      wn[it.wide].ch = std::move(ch);
      wn[it.wide].area = root.box.area();
    }
    // bottom-up sizes (children have larger indices than parents => reverse order works)
    for (int64_t i = (int64_t)wn.size() - 1; i >= 0; --i) {
      WNode& w = wn[i];
      w.own_bytes = NB; w.height = 1; w.subtree_nodes = 1;
      uint64_t sub = 0;
      for (auto& c : w.ch) {
        if (c.node < 0) w.own_bytes += (uint64_t)c.count * PRIM_BYTES;
        else { sub += wn[c.node].subtree_bytes; w.height = std::max(w.height, wn[c.node].height + 1); w.subtree_nodes += wn[c.node].subtree_nodes; }
      }
      w.subtree_bytes = w.own_bytes + sub;
    }
  }
  out->n_wide_nodes = (uint32_t)wn.size();
  lap("collapse");

  // ---- partition into treelets (BFS over treelet roots => level-contiguous ids) ----
  // Top-down greedy by surface area (the nodes a ray is most likely to visit share the subtree of their ancestors,
  // which minimises the expected number of subtree crossings = queue pushes per ray), with one rule that keeps the
  // bottom of the tree from shattering: a node whose WHOLE subtree would fit in a subtree blob of its own is never
  // split -- it is taken entirely if there is room, else it becomes an exit and later a full subtree.  (Without the
  // rule every full subtree left a fringe of 2-3 node subtrees below it: 322 K subtrees for a 10 M triangle soup.
  // A bottom-up minimum-count partition was also measured: it makes the ROOT subtree tiny, every ray is pushed
  // once per level, and cfg2 ran 6x slower.)
  struct TreeletBuild { int32_t root; uint32_t level; std::vector<int32_t> nodes; };
  std::vector<TreeletBuild> tl;
  std::vector<int32_t> node_treelet(wn.size(), -1), node_local(wn.size(), -1);
  if (n) {
    tl.push_back({0, 0, {}});
    for (size_t ti = 0; ti < tl.size(); ++ti) {
      int32_t root = tl[ti].root;
      uint32_t level = tl[ti].level;
      uint64_t used = 0;
      struct Cand { float area; int32_t node; uint32_t depth; };
      auto cmp = [](const Cand& a, const Cand& b) { return a.area < b.area; };
      std::priority_queue<Cand, std::vector<Cand>, decltype(cmp)> pq(cmp);
      std::vector<int32_t> members;
      auto include = [&](int32_t nd, uint32_t depth, auto&& self_ref) -> void {
        members.push_back(nd);
        used += wn[nd].own_bytes;
        (void)depth; (void)self_ref;
      };
      // the root is always included
      include(root, 1, include);
      for (auto& c : wn[root].ch) if (c.node >= 0) pq.push({wn[c.node].area, c.node, 2});
      std::vector<int32_t> exits;
      while (!pq.empty()) {
        Cand c = pq.top(); pq.pop();
        const WNode& w = wn[c.node];
        // whole subtree fits (bytes and stack depth): take all of it
        if (used + w.subtree_bytes <= treelet_bytes && c.depth + w.height - 1 <= depth_limit &&
            members.size() + w.subtree_nodes <= node_limit) {
          std::vector<int32_t> st{c.node};
          while (!st.empty()) {
            int32_t x = st.back(); st.pop_back();
            include(x, 0, include);
            for (auto& cc : wn[x].ch) if (cc.node >= 0) st.push_back(cc.node);
          }
          continue;
        }
        const bool fits_alone = w.subtree_bytes <= treelet_bytes && w.height <= depth_limit && w.subtree_nodes <= node_limit;
        if (!fits_alone && used + w.own_bytes <= treelet_bytes && c.depth <= depth_limit && members.size() < node_limit) {
          include(c.node, c.depth, include);
          for (auto& cc : w.ch) if (cc.node >= 0) pq.push({wn[cc.node].area, cc.node, c.depth + 1});
          continue;
        }
        exits.push_back(c.node);
      }
      // local numbering: BFS from the root over members
      for (int32_t m : members) node_treelet[m] = (int32_t)ti;
      std::vector<int32_t> order; order.reserve(members.size());
      order.push_back(root);
      for (size_t q = 0; q < order.size(); ++q)
        for (auto& cc : wn[order[q]].ch)
          if (cc.node >= 0 && node_treelet[cc.node] == (int32_t)ti) order.push_back(cc.node);
      for (size_t q = 0; q < order.size(); ++q) node_local[order[q]] = (int32_t)q;
      tl[ti].nodes = std::move(order);
      for (int32_t ex : exits) tl.push_back({ex, level + 1, {}});
      // (exits are appended in discovery order; BFS over tl keeps levels contiguous)
    }
  }
  // exits discovered from a level-L treelet have level L+1 and are appended after all level-L
  // treelets that were already queued, so ids are sorted by level.
  std::vector<int32_t> root_treelet(wn.size(), -1);
  for (size_t ti = 0; ti < tl.size(); ++ti) root_treelet[tl[ti].root] = (int32_t)ti;

  lap("treelets");
  // ---- serialise ----
  uint64_t total = 0;
  out->treelets.resize(tl.size());
  for (size_t ti = 0; ti < tl.size(); ++ti) {
    uint64_t bytes = 0;
    for (int32_t nd : tl[ti].nodes) bytes += wn[nd].own_bytes;
    bytes = (bytes + 15) & ~15ull;
    total = (total + 127) & ~127ull;
    if (total / 16 > 0xFFFFFFFFull) { set_error("BVH blob too large"); return B2RT_ERR_INVALID; }
    out->treelets[ti].offset16 = (uint32_t)(total / 16);
    out->treelets[ti].bytes = (uint32_t)bytes;
    out->treelets[ti].n_nodes = (uint32_t)tl[ti].nodes.size();
    total += bytes;
    out->max_treelet_bytes = std::max<uint32_t>(out->max_treelet_bytes, (uint32_t)bytes);
  }
  out->blob.assign((size_t)((total + 127) & ~127ull) + 128, 0);
  const float INF = std::numeric_limits<float>::infinity();
  for (size_t ti = 0; ti < tl.size(); ++ti) {
    uint8_t* base = out->blob.data() + (size_t)out->treelets[ti].offset16 * 16;
    const uint32_t nn = (uint32_t)tl[ti].nodes.size();
    uint8_t* prim_base = base + (size_t)nn * NB;
    uint32_t prim_cursor = 0;
    for (uint32_t q = 0; q < nn; ++q) {
      const WNode& w = wn[tl[ti].nodes[q]];
      float* f = reinterpret_cast<float*>(base + (size_t)q * NB);
      uint32_t* refs = reinterpret_cast<uint32_t*>(f + 6 * width);
      for (uint32_t k = 0; k < width; ++k) {
        if (k < w.ch.size()) {
          const WChild& c = w.ch[k];
          for (int a = 0; a < 3; ++a) { f[a * width + k] = c.box.mn[a]; f[(3 + a) * width + k] = c.box.mx[a]; }
          if (c.node < 0) {
            refs[k] = make_ref(REF_LEAF, ((c.count - 1) << 24) | prim_cursor);
            for (uint32_t p = 0; p < c.count; ++p) {
              uint32_t id = bb.idx[c.start + p];
              memcpy(prim_base + (size_t)(prim_cursor + p) * PRIM_BYTES, &sc.prim_geom[(size_t)id * 12], PRIM_BYTES);
            }
            prim_cursor += c.count;
            if (prim_cursor >= (1u << 24)) { set_error("too many primitives in one subtree"); return B2RT_ERR_INVALID; }
          } else if (node_treelet[c.node] == (int32_t)ti) {
            refs[k] = make_ref(REF_INTERNAL, (uint32_t)node_local[c.node]);
          } else {
            refs[k] = make_ref(REF_EXIT, (uint32_t)root_treelet[c.node]);
          }
        } else {
          for (int a = 0; a < 3; ++a) { f[a * width + k] = INF; f[(3 + a) * width + k] = -INF; }
          refs[k] = REF_EMPTY_WORD;
        }
      }
    }
    out->treelets[ti].n_prims = prim_cursor;
  }
  // levels
  uint32_t nl = 0;
  for (auto& t : tl) nl = std::max(nl, t.level + 1);
  if (nl > MAX_LEVELS) { set_error("too many subtree levels; increase treelet_bytes"); return B2RT_ERR_INVALID; }
  out->levels.assign(nl, LevelRange{0, 0});
  for (size_t ti = 0; ti < tl.size(); ++ti) {
    LevelRange& lr = out->levels[tl[ti].level];
    if (lr.count == 0) lr.first = (uint32_t)ti;
    lr.count++;
    if (ti > 0 && tl[ti].level < tl[ti - 1].level) { set_error("internal: subtree levels not sorted"); return B2RT_ERR_INVALID; }
  }
  out->n_levels = nl;
  lap("serialise");
  out->build_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
  return B2RT_OK;
}

}  // namespace b2rt

// ---- host-side structural validation of the serialised blob (no device needed) -----------------------
// Walks every subtree blob exactly as the kernel decodes it and checks: each primitive is stored in
// exactly one leaf, every child box contains the boxes of everything below it, EXIT children point
// to subtrees of the next level, blob sizes respect the shared-memory budget, local stack bound holds.
extern "C" int b2rt_bvh_validate_host(const b2rt_scene_desc* scene, uint32_t max_leaf, uint32_t width,
                                      uint32_t treelet_bytes, uint64_t out[8]) {
  using namespace b2rt;
  HostScene hs;
  int rc = make_host_scene(scene, &hs);
  if (rc) return rc;
  WideBVH wb;
  rc = build_wide_bvh(hs, max_leaf, width, treelet_bytes, &wb);
  if (rc) return rc;
  return validate_wide_bvh(hs, wb, max_leaf, treelet_bytes, out);
}

int b2rt::validate_wide_bvh(const HostScene& hs, const WideBVH& wb, uint32_t max_leaf, uint32_t treelet_bytes, uint64_t out[8]) {
  const uint32_t W = wb.width, NB = node_bytes(W);
  const uint32_t n = hs.n_prims();
  std::vector<uint32_t> seen(n, 0);
  std::vector<uint32_t> level_of(wb.treelets.size(), 0);
  for (uint32_t L = 0; L < wb.n_levels; ++L)
    for (uint32_t t = wb.levels[L].first; t < wb.levels[L].first + wb.levels[L].count; ++t) level_of[t] = L;
  uint64_t n_nodes = 0, n_leaves = 0, max_stack = 0, n_exits = 0;
  struct Bx { float mn[3], mx[3]; };
  // bounds of each subtree root (filled bottom-up: children have larger ids)
  std::vector<Bx> tl_box(wb.treelets.size());
  std::string err;
  for (int64_t t = (int64_t)wb.treelets.size() - 1; t >= 0 && err.empty(); --t) {
    const TreeletDesc& td = wb.treelets[t];
    if (td.bytes > std::max(treelet_bytes ? treelet_bytes : 24u * 1024u, NB + W * (max_leaf ? max_leaf : 4) * PRIM_BYTES)) err = "subtree exceeds byte budget";
    if ((uint64_t)td.n_nodes * NB + (uint64_t)td.n_prims * PRIM_BYTES > td.bytes) err = "subtree size mismatch";
    const uint8_t* base = wb.blob.data() + (size_t)td.offset16 * 16;
    const uint8_t* prims = base + (size_t)td.n_nodes * NB;
    // recursive bound computation over local nodes (children have larger local ids: BFS numbering)
    std::vector<Bx> nb(td.n_nodes);
    std::vector<uint32_t> depth(td.n_nodes, 1);
    for (int64_t q = (int64_t)td.n_nodes - 1; q >= 0; --q) {
      const float* f = reinterpret_cast<const float*>(base + (size_t)q * NB);
      const uint32_t* refs = reinterpret_cast<const uint32_t*>(f + 6 * W);
      Bx acc; for (int a = 0; a < 3; ++a) { acc.mn[a] = INFINITY; acc.mx[a] = -INFINITY; }
      n_nodes++;
      for (uint32_t k = 0; k < W; ++k) {
        const uint32_t r = refs[k];
        if (r == REF_EMPTY_WORD) continue;
        Bx cb; for (int a = 0; a < 3; ++a) { cb.mn[a] = f[a * W + k]; cb.mx[a] = f[(3 + a) * W + k]; }
        Bx inner; for (int a = 0; a < 3; ++a) { inner.mn[a] = INFINITY; inner.mx[a] = -INFINITY; }
        const uint32_t tag = r >> 30;
        if (tag == REF_LEAF) {
          n_leaves++;
          const uint32_t first = r & 0xFFFFFFu, cnt = ((r >> 24) & 63u) + 1;
          if (first + cnt > td.n_prims) { err = "leaf range outside subtree"; break; }
          for (uint32_t p = 0; p < cnt; ++p) {
            const float* g = reinterpret_cast<const float*>(prims + (size_t)(first + p) * PRIM_BYTES);
            uint32_t id, kind; memcpy(&id, &g[9], 4); memcpy(&kind, &g[10], 4);
            if (id >= n) { err = "primitive id out of range"; break; }
            seen[id]++;
            if (memcmp(g, &hs.prim_geom[(size_t)id * 12], PRIM_BYTES) != 0) { err = "primitive record differs from scene"; break; }
            if (kind == 0) {
              for (int v = 0; v < 3; ++v)
                for (int a = 0; a < 3; ++a) {
                  float c = g[a] + (v == 1 ? g[3 + a] : (v == 2 ? g[6 + a] : 0.f));
                  inner.mn[a] = std::min(inner.mn[a], c); inner.mx[a] = std::max(inner.mx[a], c);
                }
            } else {
              for (int a = 0; a < 3; ++a) { inner.mn[a] = std::min(inner.mn[a], g[a] - g[3]); inner.mx[a] = std::max(inner.mx[a], g[a] + g[3]); }
            }
          }
        } else if (tag == REF_INTERNAL) {
          const uint32_t c = r & 0x3FFFFFFFu;
          if (c >= td.n_nodes || c <= (uint32_t)q) { err = "bad local child index"; break; }
          inner = nb[c];
          depth[q] = std::max(depth[q], depth[c] + 1);
        } else if (tag == REF_EXIT) {
          const uint32_t c = r & 0x3FFFFFFFu;
          n_exits++;
          if (c >= wb.treelets.size() || level_of[c] != level_of[t] + 1) { err = "exit does not point to the next level"; break; }
          inner = tl_box[c];
        }
        for (int a = 0; a < 3; ++a)
          if (inner.mn[a] < cb.mn[a] || inner.mx[a] > cb.mx[a]) err = "child box does not contain its contents";
        for (int a = 0; a < 3; ++a) { acc.mn[a] = std::min(acc.mn[a], cb.mn[a]); acc.mx[a] = std::max(acc.mx[a], cb.mx[a]); }
      }
      nb[q] = acc;
    }
    if (td.n_nodes) {
      tl_box[t] = nb[0];
      max_stack = std::max<uint64_t>(max_stack, (uint64_t)(W - 1) * depth[0]);
      if (td.n_nodes > max_treelet_nodes(W)) err = "subtree has more nodes than a stack entry can address";
    }
  }
  for (uint32_t i = 0; i < n && err.empty(); ++i)
    if (seen[i] != 1) err = "primitive not stored exactly once";
  if (max_stack > stack_entries(W)) err = "per-ray stack bound exceeded";
  if (out) {
    out[0] = wb.treelets.size(); out[1] = wb.n_levels; out[2] = n_nodes; out[3] = n_leaves; out[4] = wb.blob.size();
    out[5] = wb.max_treelet_bytes; out[6] = max_stack; out[7] = n_exits;
  }
  if (!err.empty()) { set_error("bvh validation: " + err); return B2RT_ERR_INVALID; }
  return B2RT_OK;
}
