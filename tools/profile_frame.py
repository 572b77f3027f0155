#!/usr/bin/env python3
"""Short profiling workload for ncu: one frame of BASELINE configs[1] (CBbunny 1024x768, depth 8) at a reduced
sample count (--spp, default 8 = two waves).  Same kernels, same launch shapes as bench.py, fewer launches."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "cuda-raytracer_b200"))
import b2rt  # noqa: E402
import numpy as np  # noqa: E402
from b2rt.scene import Scene, place_camera, subdivide  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--spp", type=int, default=8)
ap.add_argument("--scene", default="CBbunny")
ap.add_argument("--width", type=int, default=1024)
ap.add_argument("--height", type=int, default=768)
ap.add_argument("--depth", type=int, default=8)
ap.add_argument("--bvh-width", type=int, default=0)
ap.add_argument("--treelet-bytes", type=int, default=0)
ap.add_argument("--wave", type=int, default=0)
ap.add_argument("--frames", type=int, default=1)
ap.add_argument("--max-leaf", type=int, default=0)
ap.add_argument("--overlap", action="store_true", help="no per-launch timing: the renderer overlaps the shadow and closest-hit traces on two streams")
ap.add_argument("--subdivide", type=int, default=0, help="subdivide the scene's largest mesh n times (cfg3 stand-in: CBbunny, 1)")
a = ap.parse_args()
sc = Scene.load(os.path.join(ROOT, "scenes", a.scene + ".b2s"))
if a.subdivide:
    sc = subdivide(sc, a.subdivide, select=lambda tv, tm: tm == tm[np.argmax(np.bincount(tm))])
cam = place_camera(sc, a.width, a.height)
pt = b2rt.PathTracer(ns_aa=a.spp, max_ray_depth=a.depth, ns_area_light=1, seed=1, bvh_width=a.bvh_width,
                     treelet_bytes=a.treelet_bytes, max_wave_paths=a.wave, max_leaf_size=a.max_leaf)
pt.set_scene(sc); pt.set_camera(cam); pt.set_frame_size(a.width, a.height)
pt.set_profiling(counters=False, time_kernels=not a.overlap)
for _ in range(a.frames):
    pt.clear(); pt.render()
st = pt.stats()
rays = st["rays_camera"] + st["rays_bounce"] + st["rays_shadow"]
print(f"frame {st['ms_total']:.2f} ms, traverse {st['ms_traverse']:.2f} ms, {rays / st['ms_total'] / 1e3:.1f} Mrays/s, "
      f"{st['kernel_launches']} launches ({st['traverse_launches']} traverse)")
