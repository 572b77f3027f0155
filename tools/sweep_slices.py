#!/usr/bin/env python3
"""cfg5 (10 M triangle soup, 2^24 rays): distance-slice sweep of the batch traversal (b2rt_bvh_set_slicing).
One BVH build, then (first slice in mean free paths, growth, passes) x (coherent, incoherent)."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "cuda-raytracer_b200"))
import b2rt  # noqa: E402
from b2rt.scene import random_soup  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--tris", type=int, default=10_000_000)
ap.add_argument("--rays", type=int, default=1 << 24)
ap.add_argument("--treelet-bytes", type=int, default=0)
ap.add_argument("--mfp", type=float, default=0.0, help="mean free path of the scene (0: estimate V / sum(A/2) here)")
ap.add_argument("--out", default="")
a = ap.parse_args()
soup = random_soup(a.tris)
tv = soup.tri_verts.reshape(-1, 3, 3).astype("f8")
import numpy as np  # noqa: E402
area = 0.5 * np.linalg.norm(np.cross(tv[:, 1] - tv[:, 0], tv[:, 2] - tv[:, 0]), axis=1).sum()
ext = tv.reshape(-1, 3).max(0) - tv.reshape(-1, 3).min(0)
mfp = a.mfp or float(ext.prod() / (0.5 * area))
print(f"mean free path estimate {mfp:.5f}", flush=True)
bvh = b2rt.BVHAccel(soup, treelet_bytes=a.treelet_bytes)
out = {}
for first_mfp, growth, passes in ((0, 4, 4), (-1, 4, 4), (1, 4, 5), (2, 4, 4), (3, 2, 6), (3, 4, 4), (3, 8, 3), (5, 4, 4), (8, 4, 3), (3, 4, 2)):
    bvh.set_slicing(first_mfp * mfp if first_mfp >= 0 else -1.0, growth, passes)
    for mode, name in ((0, "coherent"), (1, "incoherent")):
        ms, hits = bvh.bench_rays(a.rays, mode=mode, repeats=3)
        st = bvh.stats()
        k = f"first={first_mfp}mfp growth={growth} passes={passes} {name}"
        out[k] = dict(ms=ms, mrays_s=a.rays / ms / 1e3, hits=hits, subtree_visits_per_ray=st["subtree_visits"] / a.rays,
                      node_visits_per_ray=st["node_visits"] / a.rays, prim_tests_per_ray=st["leaf_prim_tests"] / a.rays,
                      launches=st["kernel_launches"])
        print(f"{k:55s} {ms:8.2f} ms {a.rays / ms / 1e3:8.1f} Mrays/s hits {hits} visits/ray {st['subtree_visits'] / a.rays:6.2f} "
              f"nodes/ray {st['node_visits'] / a.rays:6.1f} prims/ray {st['leaf_prim_tests'] / a.rays:6.1f}", flush=True)
if a.out:
    json.dump(out, open(a.out, "w"), indent=1)
