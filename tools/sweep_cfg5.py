#!/usr/bin/env python3
"""BASELINE configs[4]: synthetic triangle soup (10 M triangles), traversal-only ray throughput against the BVH width
W in {2, 4, 8, 16} x the leaf size in {4, 8, 16, 32} (the reference's image7.png sweeps W on a GTX 1080 Ti: 61.2 / 34.3 /
27.8 / 35.8 ms for W = 2 / 4 / 8 / 16).  Host builder (widths 2 and 16 exist only there), SAH leaf termination off so
that the leaf size is what the column says, automatic distance slicing, closest hit only.
  python tools/sweep_cfg5.py [--tris 10000000] [--rays 4194304] [--out profiles/r02_cfg5_sweep.json]"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "cuda-raytracer_b200"))
os.environ.setdefault("B2RT_SAH_CT", "1e30")
import b2rt  # noqa: E402
from b2rt.scene import random_soup  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--tris", type=int, default=10_000_000)
ap.add_argument("--rays", type=int, default=1 << 22)
ap.add_argument("--widths", default="2,4,8,16")
ap.add_argument("--leaves", default="4,8,16,32")
ap.add_argument("--out", default="")
a = ap.parse_args()
sc = random_soup(a.tris)
out = {"tris": a.tris, "rays": a.rays, "rows": []}
for W in [int(x) for x in a.widths.split(",")]:
    for leaf in [int(x) for x in a.leaves.split(",")]:
        t0 = time.time()
        try:
            bvh = b2rt.BVHAccel(sc, max_leaf_size=leaf, width=W, builder="host")
        except b2rt.B2rtError as e:
            print(f"W={W} leaf={leaf}: {e}", flush=True)
            continue
        build_s = time.time() - t0
        st = bvh.stats()
        row = dict(W=W, leaf=leaf, build_s=build_s, subtrees=st["bvh_subtrees"], levels=st["bvh_levels"], nodes=st["bvh_nodes"],
                   bvh_mb=st["bvh_bytes"] / 1e6)
        for mode, name in ((0, "coherent"), (1, "incoherent")):
            ms, hits = bvh.bench_rays(a.rays, mode=mode, repeats=3)
            s2 = bvh.stats()
            row[name] = dict(ms=ms, mrays_s=a.rays / ms / 1e3, hits=hits, node_visits_per_ray=s2["node_visits"] / a.rays,
                             prim_tests_per_ray=s2["leaf_prim_tests"] / a.rays, subtree_visits_per_ray=s2["subtree_visits"] / a.rays)
        out["rows"].append(row)
        print(f"W={W:2d} leaf={leaf:2d}: coherent {row['coherent']['mrays_s']:7.0f} Mrays/s ({row['coherent']['ms']:.2f} ms), incoherent "
              f"{row['incoherent']['mrays_s']:6.0f} Mrays/s ({row['incoherent']['ms']:.2f} ms); nodes/ray {row['coherent']['node_visits_per_ray']:.1f}/"
              f"{row['incoherent']['node_visits_per_ray']:.1f}, prims/ray {row['coherent']['prim_tests_per_ray']:.1f}/{row['incoherent']['prim_tests_per_ray']:.1f}, "
              f"{row['levels']} levels, {row['bvh_mb']:.0f} MB, build {build_s:.1f} s", flush=True)
        bvh.close()
if a.out:
    json.dump(out, open(a.out, "w"), indent=1)
