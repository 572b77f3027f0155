#!/usr/bin/env python3
"""Host (binned SAH) vs device (LBVH) BVH construction on the cfg5 soup and on the cfg3 stand-in: build time through the
C ABI (b2rt_bvh_build / b2rt_bvh_build_device, scene arrays on the host in both cases) and the ray throughput of the
resulting BVH."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "cuda-raytracer_b200"))
import b2rt  # noqa: E402
from b2rt.scene import Scene, random_soup, subdivide  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--tris", type=int, default=10_000_000)
ap.add_argument("--rays", type=int, default=1 << 24)
ap.add_argument("--validate", action="store_true")
ap.add_argument("--out", default="")
a = ap.parse_args()
out = {}
bunny = Scene.load(os.path.join(ROOT, "scenes", "CBbunny.b2s"))
scenes = [(f"soup{a.tris // 1000}K", random_soup(a.tris)),
          ("cfg3_standin_114K", subdivide(bunny, 1, select=lambda tv, tm: tm == tm[np.argmax(np.bincount(tm))]))]
for name, sc in scenes:
    for builder in ("gpu", "gpu", "host"):
        t0 = time.time()
        bvh = b2rt.BVHAccel(sc, builder=builder)
        wall = time.time() - t0
        st = bvh.stats()
        r = dict(build_wall_s=wall, build_ms_internal=st["ms_build"], subtrees=st["bvh_subtrees"], levels=st["bvh_levels"],
                 nodes=st["bvh_nodes"], bvh_mb=st["bvh_bytes"] / 1e6)
        if a.validate and builder == "gpu":
            r["validate"] = bvh.validate()
        for mode, mname in ((0, "coherent"), (1, "incoherent")):
            ms, hits = bvh.bench_rays(a.rays, mode=mode, repeats=3)
            r[mname] = dict(ms=ms, mrays_s=a.rays / ms / 1e3, hits=hits)
        out[f"{name}_{builder}"] = r
        print(name, builder, json.dumps(r), flush=True)
        bvh.close()
if a.out:
    json.dump(out, open(a.out, "w"), indent=1)
