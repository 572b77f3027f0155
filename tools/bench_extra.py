#!/usr/bin/env python3
"""Measurements beyond the headline bench line (BASELINE configs 0, 2 and 4 + BVH-width sweeps); writes
profiles/extra_r01.json.  cfg3's dragon is absent from the reference checkout (.MISSING_LARGE_BLOBS): the stand-in is
CBbunny with the bunny mesh subdivided once (114,316 triangles), SURVEY 8d.  Run on the GPU box."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "cuda-raytracer_b200"))
import b2rt  # noqa: E402
from b2rt.scene import Scene, place_camera, random_soup, subdivide  # noqa: E402

out = {}
quick = "--quick" in sys.argv
only = [a.split("=", 1)[1] for a in sys.argv if a.startswith("--only=")]   # cfg1 cfg2 cfg3 cfg5 (default: all)


def want(name):
    return not only or name in only


def frame(sc, w, h, spp, depth, nsl=1, frames=2, **kw):
    cam = place_camera(sc, w, h)
    pt = b2rt.PathTracer(ns_aa=spp, max_ray_depth=depth, ns_area_light=nsl, seed=1, **kw)
    pt.set_scene(sc); pt.set_camera(cam); pt.set_frame_size(w, h)
    pt.set_profiling(counters=False, time_kernels=True)
    for _ in range(frames):
        pt.clear(); pt.render()
    st = pt.stats()
    rays = st["rays_camera"] + st["rays_bounce"] + st["rays_shadow"]
    r = dict(ms_frame=st["ms_total"], ms_traverse=st["ms_traverse"], mrays_s=rays / st["ms_total"] / 1e3, rays=rays,
             launches=st["kernel_launches"], bvh_nodes=st["bvh_nodes"], subtrees=st["bvh_subtrees"], levels=st["bvh_levels"],
             build_ms=st["ms_build"], tris=sc.n_tris)
    pt.close()
    return r


if want("cfg1"):
    cb = Scene.load(os.path.join(ROOT, "scenes", "CBspheres_lambertian.b2s"))
    out["cfg1_CBspheres_480x360_16spp_d4"] = frame(cb, 480, 360, 16, 4)
    print("cfg1", out["cfg1_CBspheres_480x360_16spp_d4"], flush=True)

bunny = Scene.load(os.path.join(ROOT, "scenes", "CBbunny.b2s"))
for W in (4, 8) if want("cfg2") else ():
    for leaf in (2, 4, 8):
        k = f"cfg2_CBbunny_1024x768_64spp_d8_W{W}_leaf{leaf}"
        out[k] = frame(bunny, 1024, 768, 16 if quick else 64, 8, bvh_width=W, max_leaf_size=leaf)
        print(k, out[k], flush=True)

dragon = subdivide(bunny, 1, select=lambda tv, tm: tm == tm[np.argmax(np.bincount(tm))])
for W in (4, 8) if want("cfg3") else ():
    k = f"cfg3_standin_114k_1920x1080_256spp_d8_W{W}"
    out[k] = frame(dragon, 1920, 1080, 16 if quick else 256, 8, frames=1 if not quick else 2, bvh_width=W)
    print(k, out[k], flush=True)

# cfg5: traversal-only throughput on a triangle soup vs BVH width / leaf size (coherent + incoherent ray sets)
n_soup = 1_000_000 if quick else 10_000_000
soup = random_soup(n_soup) if want("cfg5") else None
for W, leaf, tb in () if soup is None else ((4, 4, 24576), (4, 4, 49152), (4, 4, 98304), (4, 8, 49152), (8, 4, 49152), (8, 8, 49152)):
    t0 = time.time()
    bvh = b2rt.BVHAccel(soup, max_leaf_size=leaf, width=W, treelet_bytes=tb)
    tb_s = time.time() - t0
    for mode, name in ((0, "coherent"), (1, "incoherent")):
        n = 1 << 22 if quick else 1 << 24
        try:
            ms, hits = bvh.bench_rays(n, mode=mode, repeats=3)
            st = bvh.stats()
            r = dict(ms=ms, mrays_s=n / ms / 1e3, hits=hits, node_visits_per_ray=st["node_visits"] / n,
                     prim_tests_per_ray=st["leaf_prim_tests"] / n, pushes_per_ray=st["queue_pushes"] / n,
                     subtree_visits_per_ray=st["subtree_visits"] / n, levels=st["bvh_levels"], subtrees=st["bvh_subtrees"],
                     bvh_mb=st["bvh_bytes"] / 1e6, build_s=tb_s)
        except b2rt.B2rtError as e:
            r = dict(error=str(e))
        k = f"cfg5_soup{n_soup // 1000000}M_W{W}_leaf{leaf}_sub{tb // 1024}K_{name}"
        out[k] = r
        print(k, r, flush=True)
    bvh.close()

os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
name = "extra_r01.json" if not only else "extra_r01_" + "_".join(only) + ".json"
json.dump(out, open(os.path.join(ROOT, "gpurun_out", name), "w"), indent=1)
