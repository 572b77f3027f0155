#!/usr/bin/env python3
"""Randomised end-to-end check on the GPU box: small random scenes (triangles with random vertex normals, spheres, all six
material kinds, area / point / directional lights, optionally an environment map) x random renderer configurations (spp,
depth, light samples, wave size, BVH width / leaf / subtree budget, builder), HDR frame compared with the oracle's.
  python tools/fuzz_frames.py [cases] [seed]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "cuda-raytracer_b200"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import b2rt  # noqa: E402
import orc  # noqa: E402
from b2rt._abi import Config  # noqa: E402
from b2rt.scene import Scene, place_camera  # noqa: E402

cases = int(sys.argv[1]) if len(sys.argv) > 1 else 24
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 20261019)
bad = 0
for case in range(cases):
    nt = int(rng.integers(20, 4000))
    c = rng.uniform(-1, 1, (nt, 1, 3)); tv = (c + rng.normal(0, rng.uniform(0.02, 0.3), (nt, 3, 3))).astype(np.float32)
    # a floor and a back wall so that paths bounce
    wall = np.array([[[-3, -1.2, -3], [3, -1.2, -3], [3, -1.2, 3]], [[-3, -1.2, -3], [3, -1.2, 3], [-3, -1.2, 3]],
                     [[-3, -1.2, -2.5], [3, 3, -2.5], [3, -1.2, -2.5]], [[-3, -1.2, -2.5], [-3, 3, -2.5], [3, 3, -2.5]]], np.float32)
    tv = np.concatenate([tv, wall])
    n = len(tv)
    normals = None
    if rng.random() < 0.5:
        e1, e2 = tv[:, 1] - tv[:, 0], tv[:, 2] - tv[:, 0]
        g = np.cross(e1, e2); g /= np.maximum(np.linalg.norm(g, axis=1, keepdims=True), 1e-20)
        normals = (g[:, None, :] + rng.normal(0, 0.2, (n, 3, 3))).astype(np.float32)
    mats = [dict(kind=0, albedo=tuple(rng.uniform(0.2, 0.9, 3))), dict(kind=1, albedo=(0.9, 0.9, 0.9)),
            dict(kind=2, albedo=(1, 1, 1), transmittance=(0.9, 0.95, 1.0), ior=1.45), dict(kind=3, emission=tuple(rng.uniform(0.5, 4, 3))),
            dict(kind=4, transmittance=(1, 1, 1), ior=1.3), dict(kind=5, albedo=(0.7, 0.7, 0.6), roughness=float(rng.uniform(0.05, 0.9))),
            dict(kind=0, albedo=(0.6, 0.6, 0.6))]
    tm = rng.choice(len(mats), n, p=[0.45, 0.08, 0.08, 0.04, 0.05, 0.1, 0.2]).astype(np.uint32)
    tm[-4:] = 6
    ns = int(rng.integers(0, 4))
    sph = np.concatenate([rng.uniform(-1, 1, (ns, 3)), rng.uniform(0.1, 0.4, (ns, 1))], 1).astype(np.float32)
    sm = rng.integers(0, len(mats), ns).astype(np.uint32)
    lights = []
    for _ in range(int(rng.integers(1, 4))):
        k = int(rng.integers(0, 3))
        if k == 0:
            lights.append(dict(kind=0, radiance=tuple(rng.uniform(2, 12, 3)), position=(float(rng.uniform(-1, 1)), 2.5, float(rng.uniform(-1, 1))),
                               direction=(0, -1, 0), dim_x=(float(rng.uniform(0.3, 1.5)), 0, 0), dim_y=(0, 0, float(rng.uniform(0.3, 1.5)))))
        elif k == 1:
            lights.append(dict(kind=1, radiance=tuple(rng.uniform(1, 6, 3)), position=tuple(rng.uniform(-1.5, 1.5, 3) + np.array([0, 1.5, 1.0]))))
        else:
            d = rng.normal(0, 1, 3); d[1] = -abs(d[1]) - 0.5; d /= np.linalg.norm(d)
            lights.append(dict(kind=2, radiance=tuple(rng.uniform(0.5, 2, 3)), direction=tuple(d)))
    sc = Scene(tv, normals, tm, sph, sm, mats, lights, cam_dir=(0, 0, -1))
    w, h = int(rng.integers(24, 97)), int(rng.integers(16, 73))
    cam = place_camera(sc, w, h)
    cfg = dict(ns_aa=int(rng.integers(1, 6)), max_ray_depth=int(rng.integers(1, 9)), ns_area_light=int(rng.integers(1, 4)),
               seed=int(rng.integers(0, 1 << 40)))
    extra = dict(bvh_width=int(rng.choice([0, 2, 4, 8, 16])), max_leaf_size=int(rng.choice([0, 1, 2, 4, 8])),
                 treelet_bytes=int(rng.choice([0, 4096, 8192, 16384, 32768])), max_wave_paths=int(rng.choice([0, 1500, 5000, 20000])),
                 bvh_builder=int(rng.choice([0, 1, 2])))
    if extra["bvh_builder"] == 2 and extra["bvh_width"] not in (0, 4, 8):
        extra["bvh_width"] = 4
    env = None
    if rng.random() < 0.3:
        ew, eh = int(rng.integers(2, 17)), int(rng.integers(2, 9))
        env = rng.uniform(0, 1.5, (eh, ew, 3)).astype(np.float32)
    pt = b2rt.PathTracer(**cfg, **extra)
    pt.set_scene(sc); pt.set_camera(cam); pt.set_frame_size(w, h)
    o = orc.OracleScene(sc, 4)
    if env is not None:
        pt.set_envmap(env); o.set_envmap(env)
    pt.render()
    img = pt.hdr()
    ref = o.render(cam, Config(**cfg), w, h)
    err = float(np.abs(img - ref).max()); rmse = float(np.sqrt(np.mean((img - ref) ** 2)))
    ok = np.isfinite(img).all() and rmse <= 1e-6 and err <= 1e-4
    bad += 0 if ok else 1
    print(f"case {case:3d}: {n} tris {ns} spheres {len(lights)} lights env {env is not None} {w}x{h} {cfg} {extra} -> max err {err:.2e} rmse {rmse:.2e} {'ok' if ok else 'MISMATCH'}", flush=True)
    pt.close()
print("mismatches:", bad)
sys.exit(1 if bad else 0)
