#!/usr/bin/env python3
"""Headline benchmark: Mrays/s (all bounces) and s/frame of the path-tracing hot path on B200.

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (one rank per GPU under torchrun)
  python bench.py --impl reference --steps K --warmup W    # CPU arm: the oracle port on the box's host cores

A "step" is one full frame of the workload.  Default = the configuration north_star states its target on, BASELINE.json
configs[2]: the dragon-class scene (stand-in: the CBbunny box with the bunny subdivided once, 114,316 triangles -- the
named CBdragon.dae is absent from the reference checkout), 1920x1080, 256 spp, max depth 8, area light, the job's 256 spp
sharded over the N ranks (STRONG scaling).  `--workload cfg2` is BASELINE configs[1] (CBbunny.dae 1024x768, 64 spp per
GPU, weak scaling; the round-1 default).  A frame = ray generation, <= 8 closest-hit traversals and <= 8 shadow-ray
traversals per path, shading, accumulation.  `value` counts every ray actually traced (camera + bounce + shadow) over
all ranks / max-over-ranks device time of the K timed steps.  Multi-GPU: samples are sharded by index (rank r renders
samples r, r+N, ...) with the scene replicated, and the per-GPU accumulation buffers are combined with ONE NCCL reduce
per frame, issued by libb2rt.so on the render stream (b2rt_reduce_accum; torch.distributed only ships the 128-byte
NCCL id and the timing scalars), inside the timed region.  `roofline` (per-launch timing of k_traverse) comes from a second pass of the same K frames with
per-launch CUDA events, in which the renderer runs on one stream (in the timed region it overlaps the shadow-ray trace
of bounce b with the closest-hit trace of bounce b + 1 on two streams, so launches have no duration of their own).
Prints exactly one JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "cuda-raytracer_b200"))

import numpy as np  # noqa: E402

WORKLOADS = {
    # name: (scene, width, height, spp, depth, ns_area_light, scaling).  "weak": spp per GPU; "strong": spp of the whole
    # job, divided over the ranks (BASELINE configs[2] / [3]: "spp sharded across 1/2/4/8 B200").  cfg3 / cfg4 use the
    # stand-in scenes of b2rt.scene (the named assets are missing from the reference checkout).  The default and the
    # driver's run is cfg3 (the scene north_star states its target on); the others are extra measurements under profiles/.
    "cfg1": ("CBspheres_lambertian", 480, 360, 16, 4, 1, "weak"),
    "cfg2": ("CBbunny", 1024, 768, 64, 8, 1, "weak"),
    "cfg3": ("cfg3_standin", 1920, 1080, 256, 8, 1, "strong"),
    "cfg4": ("cfg4_standin", 3840, 2160, 1024, 8, 1, "strong"),
}
# what the frames are made of: no data set is read; the geometry is a scene FILE of the reference checkout, converted offline
DATA_NOTE = ("synthetic workload, no data set: geometry = the reference checkout's CBbunny.dae converted to scenes/CBbunny.b2s "
             "(cfg3 / cfg4: bunny mesh subdivided once as the stand-in for the missing CBdragon / CBlucy assets); "
             "camera rays and samples generated on the device (Philox4x32-10)")
BASELINE_INDEX = {"cfg1": 0, "cfg2": 1, "cfg3": 2, "cfg4": 3}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="cfg3", choices=sorted(WORKLOADS))
    ap.add_argument("--spp", type=int, default=0, help="override samples per pixel (debug only; invalidates the number)")
    ap.add_argument("--cpu-spp", type=int, default=0, help="spp of the bounded CPU sample (default 4 of the workload's spp)")
    ap.add_argument("--bvh-width", type=int, default=0)
    ap.add_argument("--treelet-bytes", type=int, default=0)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    return ap.parse_args()


def load_workload(name, spp_override=0, world=1):
    from b2rt import scene as S
    scene_name, w, h, spp, depth, nsl, scaling = WORKLOADS[name]
    if scene_name.endswith("_standin"):
        sc = getattr(S, scene_name)(S.Scene.load(os.path.join(ROOT, "scenes", "CBbunny.b2s")))
    else:
        sc = S.Scene.load(os.path.join(ROOT, "scenes", scene_name + ".b2s"))
    cam = S.place_camera(sc, w, h)
    if spp_override:
        spp = spp_override
    if scaling == "strong":
        if spp % world:
            raise SystemExit(f"bench.py: {spp} spp do not divide over {world} ranks")
        spp //= world          # samples rank, rank + world, ... of the job's spp
    return sc, cam, dict(scene=scene_name, width=w, height=h, spp=spp, depth=depth, ns_area_light=nsl, scaling=scaling,
                         spp_job=spp * world if scaling == "strong" else spp)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md clocks line)."""

    def __init__(self, index):
        self.rows = []
        self.proc = None
        self.index = index

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                          "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.monotonic(), line.strip()))

    def stop(self, since=None):
        """Samples that arrived after `since` (the start of the timed region).  nvidia-smi needs ~100 ms to deliver its
        first line, so the sampler is started before the warm-up steps; a timed region too short to catch a sample of
        its own falls back to the samples of the warm-up steps right before it (same load) and says so."""
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        window = "timed region"
        rows = [r for t, r in self.rows if since is None or t >= since]
        if not rows:
            rows = [r for _, r in self.rows]
            window = "warm-up steps + timed region (timed region shorter than the sampling period)"
        for r in rows:
            p = [x.strip() for x in r.split(",")]
            if len(p) < 7:
                continue
            try:
                sm.append(float(p[0])); mx.append(float(p[1]))
            except ValueError:
                continue
            for n, v in zip(names, p[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons), "window": window}


def cpu_arm(args, sc, cam, wl, spp_sample, steps, warmup, target_seconds=0.0):
    """The reference's CPU implementation of the path on the host cores: the oracle port (the checkout's own
    CPU traversal / integrator bodies are stubs), all host threads, bounded sample."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import orc
    from b2rt._abi import Config
    threads = os.cpu_count() or 1
    o = orc.OracleScene(sc, 4)
    cfg = Config(ns_aa=spp_sample, max_ray_depth=wl["depth"], ns_area_light=wl["ns_area_light"], seed=1)
    for _ in range(warmup):
        o.render(cam, cfg, wl["width"], wl["height"], threads=threads, tile_stride=8)
    rays = 0
    secs = 0.0
    done = 0
    while done < steps or (target_seconds and secs < target_seconds and done < 24):
        o.render(cam, cfg, wl["width"], wl["height"], threads=threads)
        st = o.last_stats
        rays += st["rays_camera"] + st["rays_bounce"] + st["rays_shadow"]
        secs += st["seconds"]
        done += 1
    steps = done
    return dict(value=rays / secs / 1e6, unit="Mrays/s", cores=threads, kind="port",
                sample=f"{wl['scene']} {wl['width']}x{wl['height']}, {spp_sample} of {wl['spp']} spp, depth {wl['depth']}, "
                       f"{steps} frame(s), {threads} threads, oracle binary-SAH BVH (max_leaf 4)",
                seconds=secs, rays=rays, s_per_frame_scaled=secs / steps * wl["spp"] / spp_sample)


def emit(line):
    """The ONE JSON line goes to the process's original stdout (fd 1 is pointed at stderr while the bench runs, so
    that library chatter such as NCCL's version banner cannot pollute it)."""
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


_REAL_STDOUT = 1


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    sc, cam, wl = load_workload(args.workload, args.spp, world)
    metric = "Mrays/s (all bounces)"
    spp_txt = f"{wl['spp']} spp/GPU" if wl["scaling"] == "weak" else f"{wl['spp_job']} spp over {world} GPU(s)"
    config = {"workload": f"{wl['scene']}{'' if wl['scene'].endswith('_standin') else '.dae'} {wl['width']}x{wl['height']}, {spp_txt}, "
                          f"max_ray_depth {wl['depth']}, ns_area_light {wl['ns_area_light']} "
                          f"(BASELINE configs[{BASELINE_INDEX[args.workload]}])",
              "scene_tris": int(sc.n_tris), "parallelism": f"spp-sharded x{world}, scene replicated",
              "scene_source": ("generated stand-in: the bundled CBbunny.dae with its mesh midpoint-subdivided (the named asset is absent "
                               "from the reference checkout)" if wl["scene"].endswith("_standin") else "bundled .dae scene of the reference"),
              "l2": "per-wave ray/path state (<= 64Mi paths x ~200 B = GBs) exceeds the 126 MB L2; no explicit flush"}

    if args.impl == "reference":
        if rank != 0:
            return 0
        spp_s = args.cpu_spp or 4
        cb = cpu_arm(args, sc, cam, wl, spp_s, args.steps, args.warmup)
        line = {"impl": "reference", "metric": metric, "value": cb["value"], "unit": "Mrays/s", "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": cb["seconds"] / args.steps * 1e3,
                "higher_is_better": True, "scaling": wl["scaling"], "vs_baseline": None, "dtype": "f32", "data": DATA_NOTE,
                "config": config,
                "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
                "e2e": {"value": cb["value"], "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        emit(line)
        return 0

    import torch
    import b2rt
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the b200 arm has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"))
    stream = torch.cuda.current_stream()

    pt = b2rt.PathTracer(ns_aa=wl["spp"], max_ray_depth=wl["depth"], ns_area_light=wl["ns_area_light"], seed=1,
                         device=local_rank, sample_first=rank, sample_stride=world, bvh_width=args.bvh_width,
                         treelet_bytes=args.treelet_bytes)
    pt.set_stream(stream.cuda_stream)
    pt.set_scene(sc); pt.set_camera(cam); pt.set_frame_size(wl["width"], wl["height"])
    comm = None
    if world > 1:
        # the NCCL communicator lives behind the C ABI; torch.distributed only carries its 128-byte id to the ranks
        uid = [b2rt.Comm.unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        comm = b2rt.Comm(world, rank, uid[0], device=local_rank)

    def frame():
        pt.start_raytracing()     # clears the accumulation buffer, then enqueues the frame
        pt.wait()
        if comm is not None:
            pt.reduce_accum(comm, root=0)   # ONE collective per frame (ncclReduce inside libb2rt.so, on the render stream)

    # counters pass (untimed): algorithmic work of one frame
    pt.set_profiling(counters=True, time_kernels=False)
    frame()
    cst = pt.stats()
    pt.set_profiling(counters=False, time_kernels=False)
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()            # before the warm-up steps: see ClockSampler.stop
    for _ in range(args.warmup):
        frame()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t_timed = time.monotonic()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    rays = 0
    launches = 0
    ev0.record(stream)
    for _ in range(args.steps):
        frame()
        st = pt.stats()
        rays += st["rays_camera"] + st["rays_bounce"] + st["rays_shadow"]
        launches += st["kernel_launches"]
    ev1.record(stream)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    clk = clocks.stop(since=t_timed) if rank == 0 else None
    ms = ev0.elapsed_time(ev1)
    # Per-launch timing of k_traverse (roofline): the SAME K frames again with CUDA events around every traversal launch.
    # In the timed region above the renderer runs the shadow-ray trace of bounce b on a second stream next to the
    # closest-hit trace of bounce b + 1, so launches overlap and have no duration of their own; with per-launch timing on
    # the renderer uses one stream and every launch is timed alone (the same serialisation an ncu launch list shows).
    ms_trav = 0.0
    ms_trav_l0 = 0.0
    trav_launches = 0
    trav_launches_l0 = 0
    ms_iso = 0.0
    pt.set_profiling(counters=False, time_kernels=True)
    for _ in range(args.steps):
        frame()
        st = pt.stats()
        ms_trav += st["ms_traverse"]
        ms_trav_l0 += st["ms_traverse_l0"]
        trav_launches += st["traverse_launches"]
        trav_launches_l0 += st["traverse_launches_l0"]
        ms_iso += st["ms_total"]
    if world > 1:
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        r = torch.tensor([rays, launches], device="cuda", dtype=torch.float64)
        dist.all_reduce(r, op=dist.ReduceOp.SUM)
        rays_total, launches_total = int(r[0].item()), int(r[1].item())
    else:
        rays_total, launches_total = rays, launches
    value = rays_total / (ms * 1e-3) / 1e6

    # ---- end-to-end through the public API with HOST buffers: scene upload (+ host BVH build) + camera +
    #      render + read-back of the HDR frame, every step.  The scene's host arrays are page-locked (the contract's
    #      "inputs in pinned host memory"): b2rt_set_scene's copies then run as direct DMA instead of through the driver's
    #      staging buffer. ----
    pinned = []
    for name in ("tri_verts", "tri_normals", "tri_material"):
        a = getattr(sc, name)
        if a is not None and a.size:
            t = torch.from_numpy(a).pin_memory()
            pinned.append(t)
            setattr(sc, name, t.numpy())
    h2d = int(cst["bvh_bytes"] + sc.n_prims * (48 + 4) + (sc.n_tris * 36 if sc.tri_normals is not None else 0)
              + len(sc.materials) * 48 + len(sc.lights) * 64)
    d2h = wl["width"] * wl["height"] * 16      # the combined frame, read back on rank 0 only
    pt.set_profiling(counters=False, time_kernels=False)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    def e2e_step():
        pt.set_scene(sc); pt.set_camera(cam)
        pt.start_raytracing(); pt.wait()
        if comm is not None:
            pt.reduce_accum(comm, root=0)
        if rank == 0:
            pt.image()         # CudaRenderer::getImage on the root: device -> renderer-owned pinned host buffer
        else:
            torch.cuda.current_stream().synchronize()
        st = pt.stats()
        return st["rays_camera"] + st["rays_bounce"] + st["rays_shadow"]

    if args.warmup > 0:        # one untimed pass: the first getImage allocates the page-locked frame
        e2e_step()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
    t0 = time.perf_counter()
    rays_e = 0
    for _ in range(args.steps):
        rays_e += e2e_step()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
        r = torch.tensor([rays_e], device="cuda", dtype=torch.float64)
        dist.all_reduce(r, op=dist.ReduceOp.SUM)
        rays_e = int(r.item())
    e2e = {"value": rays_e / e2e_s / 1e6, "unit": "Mrays/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
           "s_per_frame": e2e_s / args.steps,
           "what": "per step and per rank: b2rt_set_scene (BVH build + upload of the scene's page-locked host arrays) + set_camera + start/wait "
                   "+ b2rt_reduce_accum; b2rt_get_image (frame to host) on rank 0"}

    if rank != 0:
        if comm is not None:
            comm.close()
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---- roofline of the dominant kernel (k_traverse) ----
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peaks = json.load(open(peaks_path)); hbm = float(peaks["hbm_gbs"]); peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        hbm = 6650.0; peak_src = "fallback (B200_PROFILING.md 6.65 TB/s)"
    W = cst["bvh_width"]
    alg_bytes = (40 * cst["subtree_visits"] + 24 * cst["queue_pushes"] + 8 * cst["hit_updates"] + cst["staged_bytes"])
    alg_flops = 24 * W * cst["node_visits"] + 55 * cst["leaf_prim_tests"]
    trav_ms_frame = ms_trav / args.steps
    tl = max(1, trav_launches // args.steps)
    achieved = alg_bytes / (trav_ms_frame * 1e-3) / 1e9
    sm_mhz = (clk or {}).get("sm_mhz") or 1965.0
    fp32_paper = 148 * 128 * 2 * sm_mhz * 1e6 / 1e12
    try:
        fp32_peak = b2rt.bench_fp32(local_rank); fp32_src = "measured on this GPU: b2rt_bench_fp32 (FFMA-saturating kernel, 16 chains per thread)"
    except b2rt.B2rtError:
        fp32_peak = fp32_paper; fp32_src = "computed (148 SMs x 128 lanes x 2 x clock under load)"
    fp32_ach = alg_flops / (trav_ms_frame * 1e-3) / 1e12
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traverse_traffic.json")
    if os.path.exists(tpath):
        traffic = (json.load(open(tpath)).get(args.workload) or {}).get("dram_bytes_per_launch")   # ncu, keyed by workload
    # per level class: level 0 streams the dense ray list (every ray once), the deeper levels gather rays by id
    rays_frame = cst["rays_camera"] + cst["rays_bounce"] + cst["rays_shadow"]
    l0 = {"visits": rays_frame, "pushes": cst["queue_pushes_l0"], "upd": cst["hit_updates_l0"], "staged": cst["staged_bytes_l0"],
          "nodes": cst["node_visits_l0"], "prims": cst["leaf_prim_tests_l0"], "ms": ms_trav_l0 / args.steps,
          "launches": max(1, trav_launches_l0 // args.steps)}
    dp = {"visits": cst["subtree_visits"] - rays_frame, "pushes": cst["queue_pushes"] - cst["queue_pushes_l0"],
          "upd": cst["hit_updates"] - cst["hit_updates_l0"], "staged": cst["staged_bytes"] - cst["staged_bytes_l0"],
          "nodes": cst["node_visits"] - cst["node_visits_l0"], "prims": cst["leaf_prim_tests"] - cst["leaf_prim_tests_l0"],
          "ms": (ms_trav - ms_trav_l0) / args.steps, "launches": max(1, (trav_launches - trav_launches_l0) // args.steps)}

    def klass(c):
        b = 40 * c["visits"] + 24 * c["pushes"] + 8 * c["upd"] + c["staged"]
        f = 24 * W * c["nodes"] + 55 * c["prims"]
        t = max(c["ms"], 1e-9) * 1e-3
        return {"ms_per_frame": c["ms"], "launches_per_step": c["launches"], "alg_bytes": b, "alg_flops": f,
                "hbm_gbs": b / t / 1e9, "hbm_frac": b / t / 1e9 / hbm, "fp32_tflops": f / t / 1e12, "fp32_frac": f / t / 1e12 / fp32_peak,
                "visits": c["visits"], "node_visits": c["nodes"], "prim_tests": c["prims"], "pushes": c["pushes"]}
    roofline = {"bound": "hbm", "achieved": achieved, "peak": hbm, "unit": "GB/s", "frac": achieved / hbm, "traffic": traffic,
                "kernel": "k_traverse", "peak_source": peak_src, "launches_per_step": tl,
                "avg_launch_ms": trav_ms_frame / tl, "alg_bytes_per_launch": alg_bytes / tl,
                "alg_bytes_per_ray": alg_bytes / max(1, rays_frame),
                "kernel_share_of_step": ms_trav / ms_iso,
                "timing": "per-launch CUDA events in a second pass of the same K frames on one stream (launches timed alone; "
                          f"that pass: {ms_iso / args.steps:.3f} ms/frame); the timed region overlaps launches on two streams",
                "fp32": {"achieved_tflops": fp32_ach, "peak_tflops": fp32_peak, "frac": fp32_ach / fp32_peak, "peak_source": fp32_src,
                         "paper_peak_tflops": fp32_paper},
                "binding_term": "hbm" if achieved / hbm >= fp32_ach / fp32_peak else "fp32",
                "by_level": {"level0": klass(l0), "deeper": klass(dp),
                             "model": "bytes = 40 x (ray, subtree) visits + 24 x pushes + 8 x hit updates + staged subtree bytes; "
                                      "flops = 24 x W x node visits + 55 x primitive tests (DESIGN.md)"},
                "counters_per_frame": {k: cst[k] for k in ("subtree_visits", "queue_pushes", "hit_updates", "staged_bytes",
                                                           "node_visits", "leaf_prim_tests")}}

    cpu_baseline = None
    if world == 1 and not args.no_cpu:
        cb = cpu_arm(args, sc, cam, wl, args.cpu_spp or 4, 2, 1, target_seconds=10.0)   # >= 2 frames, about 10 s of CPU work
        cpu_baseline = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample", "s_per_frame_scaled")}

    line = {"metric": metric, "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "s_per_frame": ms / args.steps / 1e3, "higher_is_better": True, "scaling": wl["scaling"],
            "vs_baseline": None, "dtype": "f32", "data": DATA_NOTE, "config": config, "clocks": clk, "e2e": e2e,
            "collective": ("b2rt_reduce_accum: one ncclReduce (fp32 sum) of the accumulation buffers per frame, issued by libb2rt.so "
                           f"(NCCL {b2rt.Comm.version()})" if world > 1 else None),
            "gpu_launches": launches_total, "roofline": roofline, "cpu_baseline": cpu_baseline,
            "rays_per_step": rays_total // args.steps, "bvh": {k: cst[k] for k in ("bvh_nodes", "bvh_subtrees", "bvh_levels",
                                                                                   "bvh_width", "bvh_bytes", "ms_build")}}
    emit(line)
    if comm is not None:
        comm.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
