#!/usr/bin/env python3
"""Where the end-to-end frame time of bench.py's e2e leg goes (cfg2): set_scene / set_camera / render / read-back."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "cuda-raytracer_b200"))
import b2rt  # noqa: E402
from b2rt import scene as S  # noqa: E402
from b2rt.scene import Scene, place_camera  # noqa: E402

cfg3 = len(sys.argv) > 1 and sys.argv[1] == "cfg3"          # python tools/e2e_breakdown.py [cfg3 [spp]]
sc = Scene.load(os.path.join(ROOT, "scenes", "CBbunny.b2s"))
W, H, SPP = (1920, 1080, int(sys.argv[2]) if len(sys.argv) > 2 else 32) if cfg3 else (1024, 768, 64)
if cfg3:
    sc = S.cfg3_standin(sc)
cam = place_camera(sc, W, H)
pt = b2rt.PathTracer(ns_aa=SPP, max_ray_depth=8, ns_area_light=1, seed=1)
pt.set_scene(sc); pt.set_camera(cam); pt.set_frame_size(W, H)
for _ in range(2):
    pt.start_raytracing(); pt.wait(); pt.image()
acc = {}
N = 5
for _ in range(N):
    t0 = time.perf_counter(); pt.set_scene(sc); t1 = time.perf_counter(); pt.set_camera(cam); t2 = time.perf_counter()
    pt.start_raytracing(); t3 = time.perf_counter(); pt.wait(); t4 = time.perf_counter(); img = pt.image(); t5 = time.perf_counter()
    for k, v in (("set_scene", t1 - t0), ("set_camera", t2 - t1), ("start (enqueue)", t3 - t2), ("wait", t4 - t3), ("get_image", t5 - t4),
                 ("total", t5 - t0)):
        acc[k] = acc.get(k, 0.0) + v
st = pt.stats()
for k, v in acc.items():
    print(f"{k:18s} {v / N * 1e3:8.3f} ms")
print(f"host BVH build     {st['ms_build']:8.3f} ms (inside set_scene); device frame {st['ms_total']:.3f} ms")
