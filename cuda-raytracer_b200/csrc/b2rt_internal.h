// Internal types shared by the host builder and the CUDA kernels of the b2rt library.
// Data layout in HBM is described in DESIGN.md ("Data layout").
#pragma once
#include <cstddef>
#include <cstdint>
#include <string>
#include <vector>

#include "../../include/b2rt.h"

namespace b2rt {

// ---- child reference encoding inside a wide node (32 bit) ----------------------------------------
// tag = ref >> 30 : 0 INTERNAL (payload = node index local to the subtree blob)
//                   1 LEAF     (payload = (count-1) << 24 | first primitive, local to the blob)
//                   2 EXIT     (payload = id of the child subtree the ray must be queued at)
//                   3 EMPTY
constexpr uint32_t REF_INTERNAL = 0u, REF_LEAF = 1u, REF_EXIT = 2u, REF_EMPTY = 3u;
constexpr uint32_t REF_EMPTY_WORD = 0xFFFFFFFFu;
inline uint32_t make_ref(uint32_t tag, uint32_t payload) { return (tag << 30) | payload; }

constexpr uint32_t PRIM_BYTES = 48;   // float4 x3: v0.xyz e1.x | e1.yz e2.xy | e2.z id kind pad
constexpr uint32_t MAX_LEVELS = 32;
// Per-ray traversal stack inside one subtree: 32-bit entries in shared memory, [31:12] = entry distance (fp32
// bits truncated = rounded down, so culling against it is conservative), [11:0] = local node index << log2(W) | child
// slot.  The builder bounds the subtree depth so that (W-1) * depth <= stack_entries(W), and the node count so that
// the index fits.
#ifndef B2RT_STACK4
#define B2RT_STACK4 16
#endif
#ifndef B2RT_STACK8
#define B2RT_STACK8 35
#endif
#ifndef B2RT_STACK2
#define B2RT_STACK2 16
#endif
#ifndef B2RT_STACK16
#define B2RT_STACK16 45
#endif
// widths 2 and 16 exist for the BVH-width sweep of BASELINE configs[4] (the reference's image7.png: W = 2 / 4 / 8 / 16);
// 4 is the default, 8 the alternative the device builder also produces
constexpr bool width_ok(uint32_t width) { return width == 2 || width == 4 || width == 8 || width == 16; }
constexpr uint32_t stack_entries(uint32_t width) {
  return width == 2 ? (uint32_t)B2RT_STACK2 : width == 8 ? (uint32_t)B2RT_STACK8 : width == 16 ? (uint32_t)B2RT_STACK16 : (uint32_t)B2RT_STACK4;
}
constexpr uint32_t slot_bits(uint32_t width) { return width == 2 ? 1u : width == 8 ? 3u : width == 16 ? 4u : 2u; }
constexpr uint32_t max_treelet_nodes(uint32_t width) { return 4096u >> slot_bits(width); }   // (node, slot) fits 12 bits

// Byte stride of a wide node inside a subtree blob: 32 * W bytes of rows (6 box rows + child references + 16 B spare)
// plus B2RT_NODE_PAD.  With a stride of 128 B the same row of every node falls into the same four shared-memory banks,
// so lanes of a warp that sit at DIFFERENT nodes (incoherent rays) serialise 8-fold on every 128-bit row load; a stride
// of 144 B rotates the bank group by one per node.
// Measured (tools/ab_pad.sh): 10 M soup 584 -> 605 Mrays/s incoherent, cfg2 traversal 8.99 -> 8.87 ms, cfg3 stand-in 17.28 -> 17.05 ms.
#ifndef B2RT_NODE_PAD
#define B2RT_NODE_PAD 16
#endif
constexpr uint32_t node_bytes(uint32_t width) { return 32u * width + (uint32_t)B2RT_NODE_PAD; }

struct TreeletDesc {      // one per subtree ("treelet"), 16 B
  uint32_t offset16;      // blob offset / 16
  uint32_t bytes;         // nodes + prims, multiple of 16
  uint32_t n_nodes;
  uint32_t n_prims;
};

struct LevelRange { uint32_t first, count; };

// Host-side result of the build: one contiguous blob of subtrees, position independent.
struct WideBVH {
  uint32_t width = 4;
  uint32_t n_levels = 0;
  std::vector<uint8_t> blob;
  std::vector<TreeletDesc> treelets;        // BFS order => levels are contiguous ranges
  std::vector<LevelRange> levels;
  uint32_t n_wide_nodes = 0;
  uint32_t max_treelet_bytes = 0;
  float bbox[6] = {0, 0, 0, 0, 0, 0};
  float mean_free_path = 0;             // scene-box volume / summed projected primitive area (slice length scale)
  double build_ms = 0;
};

// Per-primitive shading data in scene order (index = primitive id)
struct HostScene {
  uint32_t n_tris = 0, n_spheres = 0;
  std::vector<float> prim_geom;       // n_prims * 12 floats (PRIM_BYTES layout, scene order)
  std::vector<float> tri_normals;     // n_tris * 9 or empty
  std::vector<uint32_t> prim_material;
  std::vector<b2rt_material> materials;
  std::vector<b2rt_light> lights;
  uint32_t n_prims() const { return n_tris + n_spheres; }
};

void set_error(const std::string& msg);
// with_geometry = false: only counts, materials and lights are copied (the caller uploads tri_verts / tri_normals itself)
int make_host_scene(const b2rt_scene_desc* d, HostScene* out, bool with_geometry = true);
int build_wide_bvh(const HostScene& sc, uint32_t max_leaf, uint32_t width, uint32_t treelet_bytes, WideBVH* out);
// structural check of a serialised BVH (walks every subtree blob the way the kernel decodes it); out[8] = subtrees,
// levels, wide nodes, leaves, blob bytes, max subtree bytes, stack bound, exits
int validate_wide_bvh(const HostScene& sc, const WideBVH& bvh, uint32_t max_leaf, uint32_t treelet_bytes, uint64_t out[8]);

}  // namespace b2rt
