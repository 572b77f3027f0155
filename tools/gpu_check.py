#!/usr/bin/env python3
"""GPU bring-up check: closest-hit / any-hit parity against the oracle, render parity, quick timings.
Run on the B200 box: python tools/gpu_check.py [--quick]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "cuda-raytracer_b200"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import b2rt  # noqa: E402
import orc  # noqa: E402
from b2rt._abi import Config  # noqa: E402
from b2rt.scene import Scene, camera_rays, place_camera, random_soup  # noqa: E402


def rand_rays(scene, n, seed):
    rng = np.random.default_rng(seed)
    lo, hi = scene.bbox[:3], scene.bbox[3:]
    o = (lo + (hi - lo) * rng.random((n, 3))).astype(np.float32)
    d = rng.normal(size=(n, 3)); d /= np.linalg.norm(d, axis=1, keepdims=True)
    return o, d.astype(np.float32)


def main():
    quick = "--quick" in sys.argv
    print("devices", b2rt.device_count())
    for name, W in (("CBspheres_lambertian", 4), ("CBbunny", 4), ("CBbunny", 8), ("CBcoil", 4)):
        sc = Scene.load(os.path.join(ROOT, "scenes", name + ".b2s"))
        o = orc.OracleScene(sc, 4)
        bvh = b2rt.BVHAccel(sc, width=W)
        cam = place_camera(sc, 320, 240)
        ro, rd = camera_rays(cam, 320, 240)
        r2o, r2d = rand_rays(sc, 50000, 7)
        org = np.concatenate([ro, r2o]); dirs = np.concatenate([rd, r2d])
        t0 = time.time(); tg, pg = bvh.intersect(org, dirs); tgpu = time.time() - t0
        t0 = time.time(); tc, pc = o.intersect(org, dirs, mode="bvh"); tcpu = time.time() - t0
        bad = np.flatnonzero(pg != pc)
        badt = np.flatnonzero((tg != tc) & ~(np.isinf(tg) & np.isinf(tc)))
        print(f"{name} W{W}: rays {len(org)} prim mismatches {len(bad)} t mismatches {len(badt)} hits {np.sum(pg != 0xFFFFFFFF)} "
              f"gpu {tgpu:.3f}s cpu {tcpu:.3f}s stats {bvh.stats()}")
        if len(bad):
            for i in bad[:5]:
                print("   ray", i, "gpu", pg[i], tg[i], "cpu", pc[i], tc[i])
        # any-hit with finite tmax
        tmax = np.full(len(org), 1.5, np.float32); tmin = np.full(len(org), 1e-4, np.float32)
        occ_g = bvh.occluded(org, dirs, tmin, tmax)
        _, occ_c = o.intersect(org, dirs, tmin, tmax, any_hit=True)
        print(f"   any-hit mismatches {np.sum(occ_g != occ_c.astype(bool))} occluded {occ_g.sum()}")
        for mode in (0, 1):
            ms, hits = bvh.bench_rays(1 << 20 if quick else 1 << 22, mode=mode, repeats=3)
            st = bvh.stats()
            n = 1 << 20 if quick else 1 << 22
            print(f"   bench mode {mode}: {ms:.3f} ms -> {n / ms / 1e3:.1f} Mrays/s, hits {hits}, node_visits/ray {st['node_visits'] / n:.1f} "
                  f"prim_tests/ray {st['leaf_prim_tests'] / n:.1f} pushes/ray {st['queue_pushes'] / n:.2f} launches {st['kernel_launches']}")
        bvh.close()
    # render parity
    for name, w, h, spp, depth in (("CBspheres_lambertian", 160, 120, 4, 4), ("CBbunny", 160, 120, 4, 3), ("CBgems", 160, 120, 4, 5),
                                   ("CBcoil", 160, 120, 2, 4)):
        sc = Scene.load(os.path.join(ROOT, "scenes", name + ".b2s"))
        cam = place_camera(sc, w, h)
        pt = b2rt.PathTracer(ns_aa=spp, max_ray_depth=depth, ns_area_light=2, seed=5)
        pt.set_scene(sc); pt.set_camera(cam); pt.set_frame_size(w, h)
        t0 = time.time(); pt.render(); tg = time.time() - t0
        img = pt.hdr()
        cfg = Config(ns_aa=spp, max_ray_depth=depth, ns_area_light=2, seed=5)
        o = orc.OracleScene(sc, 4)
        ref = o.render(cam, cfg, w, h)
        diff = np.abs(img - ref)
        rmse = float(np.sqrt(np.mean((img - ref) ** 2)))
        print(f"render {name}: mean gpu {img.mean():.6f} cpu {ref.mean():.6f} rmse {rmse:.3e} max|d| {diff.max():.3e} "
              f"exact pixels {np.mean(np.all(img == ref, axis=-1)) * 100:.2f}% gpu {tg:.3f}s cpu {o.last_stats['seconds']:.3f}s")
        st = pt.stats()
        print("   gpu stats", {k: st[k] for k in ("rays_camera", "rays_bounce", "rays_shadow", "kernel_launches", "ms_total")},
              "cpu", {k: o.last_stats[k] for k in ("rays_camera", "rays_bounce", "rays_shadow")})
        pt.save_image(os.path.join(ROOT, "gpurun_out", f"{name}.png"))
        pt.close()
    if not quick:
        sc = random_soup(1000000)
        t0 = time.time(); bvh = b2rt.BVHAccel(sc); print("soup1M build", time.time() - t0, bvh.stats())
        for mode in (0, 1):
            ms, hits = bvh.bench_rays(1 << 22, mode=mode, repeats=3)
            st = bvh.stats()
            print(f"   soup bench mode {mode}: {ms:.3f} ms -> {(1 << 22) / ms / 1e3:.1f} Mrays/s hits {hits} pushes/ray {st['queue_pushes'] / (1 << 22):.2f}")


if __name__ == "__main__":
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    main()
