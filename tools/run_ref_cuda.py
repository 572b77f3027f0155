#!/usr/bin/env python3
"""Second reported baseline: the reference's own CUDA renderer (src/cudaRenderer.cu compiled unmodified for
sm_100a by oracle/build_ref.sh, driven headless by oracle/ref_cuda_driver.cpp) on the same B200, next to this repo's
renderer at the reference's operating point: 512 x 512, 2 samples per pixel per frame, 3 surface interactions with
next-event estimation at each (the reference traces a fixed script of 8 passes: 3 closest-hit + 5 shadow; here
max_ray_depth 3, ns_area_light 2 = 3 closest-hit + up to 6 shadow rays per path).  Writes gpurun_out/ref_cuda_r01.json (copied to profiles/r0N_reference_cuda.json)."""
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "cuda-raytracer_b200"))
import b2rt  # noqa: E402
from b2rt.scene import Scene, place_camera  # noqa: E402

out = {}
exe = os.path.join(ROOT, "oracle", "_ref", "ref_cuda_render")
for name in ("CBbunny", "CBcoil"):
    r = {}
    dae = os.path.join(ROOT, "oracle", "_ref", "media", name + ".dae")
    if os.path.exists(exe) and os.path.exists(dae):
        try:
            p = subprocess.run([exe, dae, "32", "3", "512"], capture_output=True, text=True, timeout=300)
            m = re.search(r"REF_CUDA_JSON (\{.*\})", p.stdout)
            r["reference_cuda"] = json.loads(m.group(1)) if m else {"error": "no result line", "rc": p.returncode, "tail": p.stdout[-400:] + p.stderr[-400:]}
        except Exception as e:  # noqa: BLE001
            r["reference_cuda"] = {"error": repr(e)}
    else:
        r["reference_cuda"] = {"error": "oracle/_ref/ref_cuda_render not built"}
    sc = Scene.load(os.path.join(ROOT, "scenes", name + ".b2s"))
    # the reference GPU renderer's view (cudaRenderer.cu:1592-1599, 347): eye = COLLADA camera position + (0, 0.75, 0)
    # = (0, 0.75, +-3) after the Z-up conversion, looking into the box, directions (+-0.5, +-0.5, 1) => 53.13 degrees both ways
    from b2rt._abi import Camera
    cam = Camera()
    sgn = 1.0 if sc.cam_dir[2] >= 0 else -1.0      # CBbunny's camera sits on the -z side, CBcoil's on the +z side
    cam.pos[:] = [0.0, 0.75, 3.0 * sgn]
    cam.c2w[:] = [sgn, 0.0, 0.0, 0.0, 1.0, 0.0, 0.0, 0.0, sgn]
    cam.hfov_deg = cam.vfov_deg = 53.13010235415598
    pt = b2rt.PathTracer(ns_aa=2, max_ray_depth=3, ns_area_light=2, seed=1)
    pt.set_scene(sc); pt.set_camera(cam); pt.set_frame_size(512, 512)
    for _ in range(4):
        pt.render()                      # CudaRenderer::render(): adds 2 spp to the running image
    import time
    t0 = time.perf_counter()
    n = 32
    rays = 0
    for _ in range(n):
        pt.render()
        st = pt.stats()
        rays += st["rays_camera"] + st["rays_bounce"] + st["rays_shadow"]
    s = time.perf_counter() - t0
    r["b2rt"] = dict(ms_per_frame=1e3 * s / n, rays_per_frame=rays / n, mrays_s=rays / s / 1e6, device_ms_last_frame=st["ms_total"])
    pt.close()
    # the same view and estimator with 64 samples per render() call: what this design is built for (large waves)
    pt = b2rt.PathTracer(ns_aa=64, max_ray_depth=3, ns_area_light=2, seed=1)
    pt.set_scene(sc); pt.set_camera(cam); pt.set_frame_size(512, 512)
    for _ in range(2):
        pt.render()
    t0 = time.perf_counter()
    rays = 0
    for _ in range(8):
        pt.render()
        st = pt.stats()
        rays += st["rays_camera"] + st["rays_bounce"] + st["rays_shadow"]
    s = time.perf_counter() - t0
    r["b2rt_64spp_per_call"] = dict(ms_per_call=1e3 * s / 8, ms_per_2spp=1e3 * s / 8 / 32, rays_per_call=rays / 8, mrays_s=rays / s / 1e6)
    pt.close()
    out[name] = r
    print(name, json.dumps(r), flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "ref_cuda_r01.json"), "w"), indent=1)
