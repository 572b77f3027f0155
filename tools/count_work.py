#!/usr/bin/env python3
"""Algorithmic work of one frame (device counters) for the library B2RT_LIB selects: node visits, primitive tests,
subtree visits and pushes per ray -- compares traversal variants (speculation cost) independent of timing."""
import argparse, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "cuda-raytracer_b200"))
import b2rt  # noqa: E402
import numpy as np  # noqa: E402
from b2rt.scene import Scene, place_camera, subdivide  # noqa: E402
ap = argparse.ArgumentParser()
ap.add_argument("--spp", type=int, default=8)
ap.add_argument("--subdivide", type=int, default=0)
ap.add_argument("--width", type=int, default=1024)
ap.add_argument("--height", type=int, default=768)
ap.add_argument("--max-leaf", type=int, default=0)
a = ap.parse_args()
sc = Scene.load(os.path.join(ROOT, "scenes", "CBbunny.b2s"))
if a.subdivide:
    sc = subdivide(sc, a.subdivide, select=lambda tv, tm: tm == tm[np.argmax(np.bincount(tm))])
cam = place_camera(sc, a.width, a.height)
pt = b2rt.PathTracer(ns_aa=a.spp, max_ray_depth=8, ns_area_light=1, seed=1, max_leaf_size=a.max_leaf)
pt.set_scene(sc); pt.set_camera(cam); pt.set_frame_size(a.width, a.height)
pt.set_profiling(counters=True, time_kernels=True)
pt.clear(); pt.render()
st = pt.stats()
rays = st["rays_camera"] + st["rays_bounce"] + st["rays_shadow"]
print(f"rays {rays/1e6:.1f}M  node_visits/ray {st['node_visits']/rays:.3f}  prim_tests/ray {st['leaf_prim_tests']/rays:.3f}  "
      f"subtree_visits/ray {st['subtree_visits']/rays:.3f}  pushes/ray {st['queue_pushes']/rays:.3f}  hit_updates/ray {st['hit_updates']/rays:.3f}  "
      f"traverse {st['ms_traverse']:.2f} ms (counters on)")
