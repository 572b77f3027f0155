// ORACLE / TEST INFRASTRUCTURE.  Driver (this repo's code) around the REFERENCE's own BVH builder:
// it is linked against src/bvh.cpp + src/bbox.cpp + CMU462 vector sources compiled from where
// they lie under /root/reference (see build_ref.sh; outputs only into oracle/_ref/).  It reads a
// .b2s scene, wraps every triangle / sphere in a Primitive whose get_bbox() restates
// Triangle::get_bbox (src/static_scene/triangle.cpp:13-47, PADDING 1e-3), runs
// BVHAccel(prims, max_leaf) + compactedTree()->compress(...) exactly as CudaRenderer::loadScene does
// (src/cudaRenderer.cu:1757-1802) and dumps the result as text for tests/test_oracle_ref.py.
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <vector>

#include "bvh.h"

using namespace CMU462;
using namespace CMU462::StaticScene;

struct FlatPrim : public Primitive {
  BBox box;
  uint32_t id;
  BBox get_bbox() const override { return box; }
  bool intersect(const Ray&) const override { return false; }
  bool intersect(const Ray&, Intersection*) const override { return false; }
  BSDF* get_bsdf() const override { return nullptr; }
  void draw(const Color&) const override {}
  void drawOutline(const Color&) const override {}
};

static void dump(BVHNode* n, FILE* f) {
  fprintf(f, "N %zu %zu %d\n", n->start, n->range, n->isLeaf() ? 1 : 0);
  if (n->l) dump(n->l, f);
  if (n->r) dump(n->r, f);
}

int main(int argc, char** argv) {
  if (argc < 4) { fprintf(stderr, "usage: ref_bvh_dump scene.b2s max_leaf out.txt\n"); return 2; }
  FILE* f = fopen(argv[1], "rb");
  if (!f) return 1;
  char magic[4]; uint32_t hdr[5]; float fh[11];
  if (fread(magic, 1, 4, f) != 4 || fread(hdr, 4, 5, f) != 5 || fread(fh, 4, 11, f) != 11) return 1;
  uint32_t nt = hdr[1], ns = hdr[2];
  std::vector<float> tv((size_t)nt * 9), tn((size_t)nt * 9), sp((size_t)ns * 4);
  std::vector<uint32_t> tm(nt), sm(ns);
  if (fread(tv.data(), 4, tv.size(), f) != tv.size()) return 1;
  if (fread(tn.data(), 4, tn.size(), f) != tn.size()) return 1;
  if (fread(tm.data(), 4, nt, f) != nt) return 1;
  if (fread(sp.data(), 4, sp.size(), f) != sp.size()) return 1;
  fclose(f);
  std::vector<Primitive*> prims;
  const double PADDING = 1e-3;
  for (uint32_t i = 0; i < nt; ++i) {
    const float* v = &tv[(size_t)i * 9];
    double mn[3], mx[3];
    for (int a = 0; a < 3; ++a) {
      double p1 = v[a], p2 = v[3 + a], p3 = v[6 + a];
      double hi = (p1 > p2) ? p1 : p2; hi = (hi > p3) ? hi : p3;
      double lo = (p1 < p2) ? p1 : p2; lo = (lo < p3) ? lo : p3;
      mn[a] = lo - PADDING; mx[a] = hi + PADDING;
    }
    FlatPrim* p = new FlatPrim();
    p->box = BBox(mn[0], mn[1], mn[2], mx[0], mx[1], mx[2]);
    p->id = i;
    prims.push_back(p);
  }
  for (uint32_t i = 0; i < ns; ++i) {
    const float* s = &sp[(size_t)i * 4];
    FlatPrim* p = new FlatPrim();
    p->box = BBox((double)s[0] - (double)s[3], (double)s[1] - (double)s[3], (double)s[2] - (double)s[3],
                  (double)s[0] + (double)s[3], (double)s[1] + (double)s[3], (double)s[2] + (double)s[3]);
    p->id = nt + i;
    prims.push_back(p);
  }
  size_t max_leaf = (size_t)atoi(argv[2]);
  BVHAccel* bvh = new BVHAccel(prims, max_leaf);
  FILE* o = fopen(argv[3], "w");
  fprintf(o, "PRIMS %zu\n", prims.size());
  std::vector<Primitive*> sorted = bvh->getSortedPrimitives();
  for (Primitive* p : sorted) fprintf(o, "P %u\n", static_cast<FlatPrim*>(p)->id);
  dump(bvh->get_root(), o);
  // wide collapse as loadScene does it: LEVEL_INDEX_SIZE 6000, MAX_LEVELS 16 (cudaRenderer.h:63-64);
  // the level lists are sized generously here so large scenes do not overflow the reference's arrays.
  const int stride = 400000, levels = 64;
  std::vector<int> levelIndices((size_t)stride * levels);
  std::vector<int> levelCounts;
  std::vector<C_BVHSubTree> tree;
  bvh->compactedTree()->compress(&tree, levelIndices.data(), stride, &levelCounts, 0, levels - 2);
  fprintf(o, "WIDE %zu\n", tree.size());
  for (int c : levelCounts) fprintf(o, "L %d\n", c);
  fclose(o);
  return 0;
}
