#!/usr/bin/env python3
"""Experiment: two renderer handles on ONE device, each rendering half of the samples (sample_first 0 / 1, stride 2) from
two host threads, vs one handle rendering all 64 spp: does a deeper overlap (4 streams) gain over the 2-stream renderer?"""
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "cuda-raytracer_b200"))
import b2rt  # noqa: E402
from b2rt.scene import Scene, place_camera  # noqa: E402

sc = Scene.load(os.path.join(ROOT, "scenes", "CBbunny.b2s"))
cam = place_camera(sc, 1024, 768)


def make(spp, first, stride, wave=0):
    pt = b2rt.PathTracer(ns_aa=spp, max_ray_depth=8, ns_area_light=1, seed=1, sample_first=first, sample_stride=stride, max_wave_paths=wave)
    pt.set_scene(sc); pt.set_camera(cam); pt.set_frame_size(1024, 768)
    return pt


def run(pts, n=5):
    for pt in pts:
        pt.clear(); pt.start_raytracing()
    for pt in pts:
        pt.wait()
    t0 = time.perf_counter()
    for _ in range(n):
        for pt in pts:
            pt.clear(); pt.start_raytracing()
        for pt in pts:
            pt.wait()
    return (time.perf_counter() - t0) / n * 1e3


one = make(64, 0, 1)
print(f"one handle, 64 spp            : {run([one]):.2f} ms/frame")
one.close()
two = [make(32, 0, 2), make(32, 1, 2)]
print(f"two handles, 32 spp each      : {run(two):.2f} ms/frame")
for p in two:
    p.close()
four = [make(16, k, 4) for k in range(4)]
print(f"four handles, 16 spp each     : {run(four):.2f} ms/frame")
