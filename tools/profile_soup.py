#!/usr/bin/env python3
"""Short profiling workload for ncu on BASELINE configs[4]: closest-hit traversal of a synthetic random triangle soup
(10 M triangles, size 0.01, seed 0x5EED0001) with 2^24 coherent or incoherent rays."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "cuda-raytracer_b200"))
import b2rt  # noqa: E402
from b2rt.scene import random_soup  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--tris", type=int, default=10_000_000)
ap.add_argument("--rays", type=int, default=1 << 24)
ap.add_argument("--mode", type=int, default=0, help="0 coherent, 1 incoherent")
ap.add_argument("--bvh-width", type=int, default=4)
ap.add_argument("--max-leaf", type=int, default=4)
ap.add_argument("--treelet-bytes", type=int, default=0)
ap.add_argument("--repeats", type=int, default=1)
ap.add_argument("--builder", default="host", choices=["host", "gpu"])
a = ap.parse_args()
soup = random_soup(a.tris)
bvh = b2rt.BVHAccel(soup, max_leaf_size=a.max_leaf, width=a.bvh_width, treelet_bytes=a.treelet_bytes, builder=a.builder)
ms, hits = bvh.bench_rays(a.rays, mode=a.mode, repeats=a.repeats)
st = bvh.stats()
print(f"soup {a.tris} tris, {a.rays} rays mode {a.mode}: {ms:.2f} ms, {a.rays / ms / 1e3:.1f} Mrays/s, {st['bvh_subtrees']} subtrees in "
      f"{st['bvh_levels']} levels, {st['subtree_visits'] / a.rays:.1f} subtree visits/ray, {st['node_visits'] / a.rays:.1f} node visits/ray")
