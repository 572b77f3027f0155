#!/usr/bin/env python3
"""Device BVH build of the cfg5 soup, for an ncu launch list (the second build runs with a warm memory pool)."""
import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "cuda-raytracer_b200"))
import b2rt  # noqa: E402
from b2rt.scene import random_soup  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--tris", type=int, default=10_000_000)
ap.add_argument("--builds", type=int, default=2)
a = ap.parse_args()
soup = random_soup(a.tris)
for k in range(a.builds):
    t0 = time.time()
    bvh = b2rt.BVHAccel(soup, builder="gpu")
    st = bvh.stats()
    print(f"build {k}: {st['ms_build']:.2f} ms inside the library, {1e3 * (time.time() - t0):.1f} ms wall, {st['bvh_subtrees']} subtrees, "
          f"{st['bvh_levels']} levels, {st['bvh_nodes']} wide nodes, {st['bvh_bytes'] / 1e6:.1f} MB")
    bvh.close()
