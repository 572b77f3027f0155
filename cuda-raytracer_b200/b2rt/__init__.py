"""Python host mirror of the reference's renderer interface for the B200 hot path.

The product is `libb2rt.so` (C ABI, include/b2rt.h; CUDA for sm_100a only).  This module is a thin
ctypes binding whose classes keep the reference's names and call order:

  PathTracer   <-> class PathTracer     (src/pathtracer.h:51-257)
  BVHAccel     <-> class BVHAccel       (src/bvh.h:99-191), batch intersect instead of per-ray virtual calls
  CudaRenderer <-> class CudaRenderer   (src/cudaRenderer.h:173-272), progressive render()/getImage()

There is NO CPU fallback: if the library is missing or no CUDA device is visible, device calls raise.
"""
import ctypes as C
import os

import numpy as np

from ._abi import Camera, Config, Light, Material, SceneDesc, SceneFile, Stats
from .scene import Scene, camera_rays, place_camera, random_soup, subdivide  # noqa: F401

_HERE = os.path.dirname(os.path.abspath(__file__))
# B2RT_LIB selects another build of the same library (A/B runs of kernel variants, tools/build_variant.sh)
LIB_PATH = os.environ.get("B2RT_LIB") or os.path.join(os.path.dirname(_HERE), "libb2rt.so")

EXPORTS = [
    "b2rt_last_error", "b2rt_abi_version", "b2rt_device_count", "b2rt_bvh_build", "b2rt_bvh_build_device", "b2rt_bvh_validate",
    "b2rt_bvh_intersect",
    "b2rt_bvh_occluded", "b2rt_bvh_bench_rays", "b2rt_bvh_set_slicing", "b2rt_bvh_get_stats", "b2rt_bvh_get_bbox", "b2rt_bvh_destroy",
    "b2rt_create", "b2rt_set_config", "b2rt_set_scene", "b2rt_set_camera", "b2rt_set_frame_size", "b2rt_start",
    "b2rt_is_done", "b2rt_wait", "b2rt_stop", "b2rt_clear", "b2rt_render", "b2rt_read_hdr", "b2rt_read_ldr",
    "b2rt_read_rgba32f", "b2rt_get_image", "b2rt_get_stats", "b2rt_accum_device_ptr", "b2rt_stream_handle", "b2rt_set_stream",
    "b2rt_set_profiling", "b2rt_destroy", "b2rt_bvh_validate_host",
    "b2rt_scene_load", "b2rt_scene_save", "b2rt_load_dae", "b2rt_scene_free", "b2rt_camera_place",
    "b2rt_comm_version", "b2rt_comm_unique_id", "b2rt_comm_create", "b2rt_comm_create_all", "b2rt_reduce_accum",
    "b2rt_reduce_accum_all", "b2rt_comm_destroy", "b2rt_bench_fp32", "b2rt_camera_look_at", "b2rt_save_png", "b2rt_save_exr",
    "b2rt_write_png", "b2rt_write_exr", "b2rt_set_envmap",
]


class B2rtError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"b2rt error {code}: {msg}")
        self.code = code


_lib = None


def lib():
    """Load libb2rt.so (fails loudly when it has not been built: python __graft_entry__.py build)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise FileNotFoundError(f"{LIB_PATH} not built; run `make -C cuda-raytracer_b200/csrc` "
                                    "(or __graft_entry__.build()); there is no CPU fallback")
        L = C.CDLL(LIB_PATH)
        L.b2rt_last_error.restype = C.c_char_p
        vp, u32, u64, i32 = C.c_void_p, C.c_uint32, C.c_uint64, C.c_int32
        L.b2rt_bvh_build.argtypes = [C.POINTER(SceneDesc), u32, u32, u32, i32, C.POINTER(vp)]
        L.b2rt_bvh_build_device.argtypes = [C.POINTER(SceneDesc), u32, u32, u32, i32, C.POINTER(vp)]
        L.b2rt_bvh_validate.argtypes = [vp, C.POINTER(SceneDesc), vp]
        L.b2rt_bvh_intersect.argtypes = [vp, vp, vp, vp, vp, u64, vp, vp]
        L.b2rt_bvh_occluded.argtypes = [vp, vp, vp, vp, vp, u64, vp]
        L.b2rt_bvh_bench_rays.argtypes = [vp, u64, C.c_int, u64, C.c_int, C.c_int, C.POINTER(C.c_double), C.POINTER(u64)]
        L.b2rt_bvh_set_slicing.argtypes = [vp, C.c_float, C.c_float, i32]
        L.b2rt_bvh_get_stats.argtypes = [vp, C.POINTER(Stats)]
        L.b2rt_bvh_get_bbox.argtypes = [vp, vp]
        L.b2rt_bvh_destroy.argtypes = [vp]; L.b2rt_bvh_destroy.restype = None
        L.b2rt_create.argtypes = [C.POINTER(Config), C.POINTER(vp)]
        L.b2rt_set_config.argtypes = [vp, C.POINTER(Config)]
        L.b2rt_set_scene.argtypes = [vp, C.POINTER(SceneDesc)]
        L.b2rt_set_camera.argtypes = [vp, C.POINTER(Camera)]
        L.b2rt_set_frame_size.argtypes = [vp, u32, u32]
        for f in ("b2rt_start", "b2rt_is_done", "b2rt_wait", "b2rt_stop", "b2rt_clear", "b2rt_render"):
            getattr(L, f).argtypes = [vp]
        L.b2rt_read_hdr.argtypes = [vp, vp, C.c_size_t]
        L.b2rt_read_ldr.argtypes = [vp, vp, C.c_size_t]
        L.b2rt_read_rgba32f.argtypes = [vp, vp, C.c_size_t]
        L.b2rt_get_image.argtypes = [vp, C.POINTER(C.POINTER(C.c_float)), C.POINTER(C.c_size_t)]
        L.b2rt_get_stats.argtypes = [vp, C.POINTER(Stats)]
        L.b2rt_accum_device_ptr.argtypes = [vp, C.POINTER(vp), C.POINTER(C.c_size_t)]
        L.b2rt_stream_handle.argtypes = [vp, C.POINTER(vp)]
        L.b2rt_set_stream.argtypes = [vp, vp]
        L.b2rt_set_profiling.argtypes = [vp, C.c_int, C.c_int]
        L.b2rt_bvh_validate_host.argtypes = [C.POINTER(SceneDesc), u32, u32, u32, vp]
        L.b2rt_destroy.argtypes = [vp]; L.b2rt_destroy.restype = None
        L.b2rt_camera_place.argtypes = [vp, vp, C.c_float, C.c_float, u32, u32, C.POINTER(Camera)]
        for f in ("b2rt_scene_load", "b2rt_load_dae"):
            getattr(L, f).argtypes = [C.c_char_p, C.POINTER(vp)]
        L.b2rt_scene_save.argtypes = [C.c_char_p, vp]
        L.b2rt_scene_free.argtypes = [vp]; L.b2rt_scene_free.restype = None
        L.b2rt_bench_fp32.argtypes = [i32, C.POINTER(C.c_double)]
        L.b2rt_camera_look_at.argtypes = [vp, vp, C.c_float, C.POINTER(Camera)]
        L.b2rt_save_png.argtypes = [C.c_char_p, vp, u32, u32]
        L.b2rt_save_exr.argtypes = [C.c_char_p, vp, u32, u32]
        L.b2rt_write_png.argtypes = [vp, C.c_char_p]
        L.b2rt_write_exr.argtypes = [vp, C.c_char_p]
        L.b2rt_set_envmap.argtypes = [vp, vp, u32, u32]
        L.b2rt_comm_unique_id.argtypes = [vp]
        L.b2rt_comm_create.argtypes = [i32, i32, vp, i32, C.POINTER(vp)]
        L.b2rt_comm_create_all.argtypes = [i32, vp, vp]
        L.b2rt_reduce_accum.argtypes = [vp, vp, i32]
        L.b2rt_reduce_accum_all.argtypes = [vp, vp, i32, i32]
        L.b2rt_comm_destroy.argtypes = [vp]; L.b2rt_comm_destroy.restype = None
        _lib = L
    return _lib


def _check(rc):
    if rc < 0:
        raise B2rtError(rc, lib().b2rt_last_error().decode())
    return rc


def device_count():
    return lib().b2rt_device_count()


def camera_look_at(origin, look_at, fov_deg=0.0):
    """Camera of CudaRenderer::setViewpoint(origin, lookAt) (src/cudaRenderer.cu:1845-1870), b2rt_camera_look_at."""
    o = np.ascontiguousarray(origin, np.float32); d = np.ascontiguousarray(look_at, np.float32)
    cam = Camera()
    _check(lib().b2rt_camera_look_at(o.ctypes.data, d.ctypes.data, fov_deg, C.byref(cam)))
    return cam


def save_png(path, rgba8):
    """rgba8: uint32 [h, w] as PathTracer.ldr() returns it (row 0 = bottom); the file stores the top row first."""
    a = np.ascontiguousarray(rgba8, np.uint32)
    _check(lib().b2rt_save_png(os.fsencode(path), a.ctypes.data, a.shape[1], a.shape[0]))


def save_exr(path, rgb):
    """rgb: float32 [h, w, 3] as PathTracer.hdr() returns it; uncompressed fp32 OpenEXR scanline file."""
    a = np.ascontiguousarray(rgb, np.float32)
    _check(lib().b2rt_save_exr(os.fsencode(path), a.ctypes.data, a.shape[1], a.shape[0]))


def bench_fp32(device=-1):
    """Measured FP32 peak of the device in TFLOP/s (b2rt_bench_fp32)."""
    v = C.c_double(0)
    _check(lib().b2rt_bench_fp32(device, C.byref(v)))
    return v.value


def _scene_from_file(entry, path):
    """Run b2rt_scene_load / b2rt_load_dae and copy the result into a host Scene (the library frees its copy)."""
    pf = C.c_void_p()
    _check(getattr(lib(), entry)(os.fsencode(path), C.byref(pf)))
    try:
        f = C.cast(pf, C.POINTER(SceneFile)).contents
        d = f.desc
        def arr(ptr, n, dt):
            return np.ctypeslib.as_array(ptr, shape=(n,)).astype(dt, copy=True) if n and ptr else np.zeros(0, dt)
        mats = [dict(kind=m.kind, albedo=tuple(m.albedo), transmittance=tuple(m.transmittance), emission=tuple(m.emission),
                     ior=m.ior, roughness=m.roughness) for m in (d.materials[i] for i in range(d.n_materials))]
        lights = [dict(kind=l.kind, radiance=tuple(l.radiance), position=tuple(l.position), direction=tuple(l.direction),
                       dim_x=tuple(l.dim_x), dim_y=tuple(l.dim_y)) for l in (d.lights[i] for i in range(d.n_lights))]
        return Scene(arr(d.tri_verts, d.n_tris * 9, np.float32),
                     arr(d.tri_normals, d.n_tris * 9, np.float32) if d.tri_normals else None,
                     arr(d.tri_material, d.n_tris, np.uint32), arr(d.spheres, d.n_spheres * 4, np.float32),
                     arr(d.sphere_material, d.n_spheres, np.uint32), mats, lights, cam_dir=tuple(f.cam_dir),
                     hfov=f.cam_hfov_deg, vfov=f.cam_vfov_deg, bbox=np.array(tuple(f.bbox), np.float64))
    finally:
        lib().b2rt_scene_free(pf)


def load_dae(path):
    """Collada::ColladaParser::load + Application::load flattening (src/collada/collada.cpp:117-214,
    src/application.cpp:347-435) through the C ABI's b2rt_load_dae -> Scene."""
    return _scene_from_file("b2rt_load_dae", path)


def load_scene(path):
    """.dae through b2rt_load_dae, anything else as a .b2s file through b2rt_scene_load."""
    return _scene_from_file("b2rt_load_dae" if str(path).lower().endswith(".dae") else "b2rt_scene_load", path)


def _f32(a, shape=None):
    a = np.ascontiguousarray(a, np.float32)
    return a if shape is None else a.reshape(shape)


def validate_bvh_host(scene, max_leaf_size=4, width=4, treelet_bytes=0):
    """Host-only structural check of the serialised BVH (no GPU)."""
    d, keep = scene.desc()
    out = np.zeros(8, np.uint64)
    _check(lib().b2rt_bvh_validate_host(C.byref(d), max_leaf_size, width, treelet_bytes, out.ctypes.data))
    keys = ("subtrees", "levels", "wide_nodes", "leaves", "blob_bytes", "max_subtree_bytes", "stack_bound", "exits")
    return dict(zip(keys, out.tolist()))


class BVHAccel:
    """BVHAccel(primitives, max_leaf_size) -- src/bvh.h:110-112; closest / any-hit queries are batched."""

    def __init__(self, scene, max_leaf_size=4, width=4, treelet_bytes=0, device=-1, builder="host"):
        """builder: "host" = binned-SAH build on the host cores (b2rt_bvh_build), "gpu" = LBVH build on the device
        (b2rt_bvh_build_device)."""
        d, keep = scene.desc()
        h = C.c_void_p()
        fn = {"host": lib().b2rt_bvh_build, "gpu": lib().b2rt_bvh_build_device}[builder]
        _check(fn(C.byref(d), max_leaf_size, width, treelet_bytes, device, C.byref(h)))
        self._h = h
        self.scene = scene

    def validate(self):
        """Structural check of the device-resident BVH (b2rt_bvh_validate); returns the counts."""
        d, keep = self.scene.desc()
        out = np.zeros(8, np.uint64)
        _check(lib().b2rt_bvh_validate(self._h, C.byref(d), out.ctypes.data))
        keys = ("subtrees", "levels", "wide_nodes", "leaves", "blob_bytes", "max_subtree_bytes", "stack_bound", "exits")
        return dict(zip(keys, (int(v) for v in out)))

    def close(self):
        if getattr(self, "_h", None):
            lib().b2rt_bvh_destroy(self._h); self._h = None

    __del__ = close

    def get_bbox(self):
        out = np.zeros(6, np.float32)
        _check(lib().b2rt_bvh_get_bbox(self._h, out.ctypes.data))
        return out

    def _rays(self, org, dirs, tmin, tmax):
        org = _f32(org, (-1, 3)); dirs = _f32(dirs, (-1, 3))
        n = len(org)
        tmin = np.zeros(n, np.float32) if tmin is None else _f32(tmin)
        tmax = np.full(n, np.inf, np.float32) if tmax is None else _f32(tmax)
        return org, dirs, tmin, tmax, n

    def intersect(self, org, dirs, tmin=None, tmax=None):
        """Closest hit: returns (t[n] float32 (inf on miss), prim[n] uint32 (0xFFFFFFFF on miss))."""
        org, dirs, tmin, tmax, n = self._rays(org, dirs, tmin, tmax)
        t = np.empty(n, np.float32); prim = np.empty(n, np.uint32)
        _check(lib().b2rt_bvh_intersect(self._h, org.ctypes.data, dirs.ctypes.data, tmin.ctypes.data, tmax.ctypes.data, n,
                                        t.ctypes.data, prim.ctypes.data))
        return t, prim

    def occluded(self, org, dirs, tmin=None, tmax=None):
        org, dirs, tmin, tmax, n = self._rays(org, dirs, tmin, tmax)
        occ = np.empty(n, np.uint8)
        _check(lib().b2rt_bvh_occluded(self._h, org.ctypes.data, dirs.ctypes.data, tmin.ctypes.data, tmax.ctypes.data, n,
                                       occ.ctypes.data))
        return occ.astype(bool)

    def bench_rays(self, n, mode=0, seed=1, repeats=3, any_hit=False):
        ms = C.c_double(0); hits = C.c_uint64(0)
        _check(lib().b2rt_bvh_bench_rays(self._h, n, mode, seed, repeats, int(any_hit), C.byref(ms), C.byref(hits)))
        return ms.value, hits.value

    def set_slicing(self, first_slice=-1.0, growth=4.0, passes=4):
        """Distance slices of the batch traversal (b2rt_bvh_set_slicing): > 0 explicit first slice, 0 off, < 0 automatic."""
        _check(lib().b2rt_bvh_set_slicing(self._h, first_slice, growth, passes))

    def stats(self):
        s = Stats()
        _check(lib().b2rt_bvh_get_stats(self._h, C.byref(s)))
        return s.as_dict()


class PathTracer:
    """Mirror of class PathTracer (src/pathtracer.h:51-257): same knobs, same call order.

    pt = PathTracer(ns_aa=16, max_ray_depth=4, ns_area_light=1); pt.set_scene(scene); pt.set_camera(cam)
    pt.set_frame_size(w, h); pt.start_raytracing(); while not pt.is_done(): ...; pt.save_image("out.png")
    """
    INIT, READY, RENDERING, DONE = "INIT", "READY", "RENDERING", "DONE"

    def __init__(self, ns_aa=1, max_ray_depth=4, ns_area_light=1, ns_diff=1, ns_glsy=1, ns_refr=1, num_threads=1,
                 envmap=None, seed=0, device=-1, bvh_width=0, max_leaf_size=0, treelet_bytes=0, max_wave_paths=0,
                 median_threshold=0, sample_first=0, sample_stride=1, ray_eps=0.0, bvh_builder=0, filter_kind=0,
                 filter_sigma_r=0.0):
        self.cfg = Config(ns_aa=ns_aa, max_ray_depth=max_ray_depth, ns_area_light=ns_area_light, seed=seed,
                          ray_eps=ray_eps, bvh_width=bvh_width, max_leaf_size=max_leaf_size,
                          treelet_bytes=treelet_bytes, max_wave_paths=max_wave_paths,
                          median_threshold=median_threshold, device=device, sample_first=sample_first,
                          sample_stride=sample_stride, bvh_builder=bvh_builder, filter_kind=filter_kind,
                          filter_sigma_r=filter_sigma_r)
        h = C.c_void_p()
        _check(lib().b2rt_create(C.byref(self.cfg), C.byref(h)))
        self._h = h
        self._scene = self._camera = None
        self.width = self.height = 0
        self.state = self.INIT
        if envmap is not None:
            self.set_envmap(envmap)

    def close(self):
        if getattr(self, "_h", None):
            lib().b2rt_destroy(self._h); self._h = None

    __del__ = close

    def _maybe_ready(self):
        if self.state == self.INIT and self._scene is not None and self._camera is not None and self.width:
            self.state = self.READY

    def set_config(self, **kw):
        for k, v in kw.items():
            setattr(self.cfg, k, v)
        _check(lib().b2rt_set_config(self._h, C.byref(self.cfg)))

    def set_scene(self, scene):
        d, keep = scene.desc()
        _check(lib().b2rt_set_scene(self._h, C.byref(d)))
        self._scene = scene
        if self.state != self.INIT:
            self.state = self.READY
        self._maybe_ready()

    def set_camera(self, camera):
        _check(lib().b2rt_set_camera(self._h, C.byref(camera)))
        self._camera = camera
        if self.state != self.INIT:
            self.state = self.READY
        self._maybe_ready()

    def set_envmap(self, rgb):
        """Environment light (the `envmap` argument of PathTracer::PathTracer, src/pathtracer.h:57-60): float32 [h, w, 3],
        row 0 = the +y pole, x = azimuth atan2(d.z, d.x) / 2 pi; None removes it."""
        if rgb is None:
            _check(lib().b2rt_set_envmap(self._h, None, 0, 0))
        else:
            a = np.ascontiguousarray(rgb, np.float32)
            _check(lib().b2rt_set_envmap(self._h, a.ctypes.data, a.shape[1], a.shape[0]))

    def set_frame_size(self, width, height):
        _check(lib().b2rt_set_frame_size(self._h, width, height))
        self.width, self.height = width, height
        if self.state != self.INIT:
            self.state = self.READY      # pathtracer.cpp:105-114
        self._maybe_ready()

    def start_raytracing(self):
        """PathTracer::start_raytracing (src/pathtracer.cpp:183-213): only from READY (DONE re-renders); like the
        reference it clears the sample / frame buffers first.  render() is the accumulating call."""
        if self.state not in (self.READY, self.DONE):
            return False
        _check(lib().b2rt_clear(self._h))
        _check(lib().b2rt_start(self._h))
        self.state = self.RENDERING
        return True

    def is_done(self):
        if self.state == self.RENDERING and _check(lib().b2rt_is_done(self._h)) == 1:
            self.state = self.DONE
        return self.state == self.DONE

    def wait(self):
        _check(lib().b2rt_wait(self._h))
        if self.state == self.RENDERING:
            self.state = self.DONE

    def stop(self):
        _check(lib().b2rt_stop(self._h))
        if self.state in (self.RENDERING, self.DONE):
            self.state = self.READY

    def clear(self):
        _check(lib().b2rt_clear(self._h))

    def render(self):
        """Blocking start + wait (accumulates ns_aa more samples, like CudaRenderer::render)."""
        _check(lib().b2rt_render(self._h))
        self.state = self.DONE

    def increase_area_light_sample_count(self):      # pathtracer.cpp:560-564
        self.set_config(ns_area_light=self.cfg.ns_area_light * 2)

    def decrease_area_light_sample_count(self):      # pathtracer.cpp:566-570
        self.set_config(ns_area_light=max(1, self.cfg.ns_area_light // 2))

    def hdr(self):
        """HDRImageBuffer: float32 [h, w, 3], row 0 = bottom row, index x + y*w (src/image.h:114-118)."""
        out = np.empty((self.height, self.width, 3), np.float32)
        _check(lib().b2rt_read_hdr(self._h, out.ctypes.data, out.size))
        return out

    def ldr(self):
        """ImageBuffer: uint32 RGBA8 [h, w] via toColor (src/image.h:49-58,173-188)."""
        out = np.empty((self.height, self.width), np.uint32)
        _check(lib().b2rt_read_ldr(self._h, out.ctypes.data, out.size))
        return out

    def rgba32f(self):
        out = np.empty((self.height, self.width, 4), np.float32)
        _check(lib().b2rt_read_rgba32f(self._h, out.ctypes.data, out.size))
        return out

    def image(self):
        """CudaRenderer::getImage: float4 RGBA view (height, width, 4) of the renderer-owned page-locked host buffer;
        valid until the next call on this object (copy it to keep it)."""
        p = C.POINTER(C.c_float)(); n = C.c_size_t(0)
        _check(lib().b2rt_get_image(self._h, C.byref(p), C.byref(n)))
        return np.ctypeslib.as_array(p, shape=(self.height, self.width, 4))

    def save_image(self, filename):
        """PNG, vertically flipped like PathTracer::save_image (src/pathtracer.cpp:577-591); written by the library."""
        _check(lib().b2rt_write_png(self._h, os.fsencode(filename)))

    def save_exr(self, filename):
        _check(lib().b2rt_write_exr(self._h, os.fsencode(filename)))

    def stats(self):
        s = Stats()
        _check(lib().b2rt_get_stats(self._h, C.byref(s)))
        return s.as_dict()

    def set_stream(self, cuda_stream_ptr):
        _check(lib().b2rt_set_stream(self._h, C.c_void_p(cuda_stream_ptr)))

    def set_profiling(self, counters=False, time_kernels=False):
        _check(lib().b2rt_set_profiling(self._h, int(counters), int(time_kernels)))

    def reduce_accum(self, comm, root=0):
        """ONE NCCL reduce (fp32 sum) of the per-GPU accumulation buffers onto `root`, enqueued by libb2rt.so on the
        renderer's stream (b2rt_reduce_accum); root < 0 = all-reduce.  `comm` is a b2rt.Comm."""
        _check(lib().b2rt_reduce_accum(self._h, comm._h, root))

    def accum_device_ptr(self):
        p = C.c_void_p(); n = C.c_size_t(0)
        _check(lib().b2rt_accum_device_ptr(self._h, C.byref(p), C.byref(n)))
        return p.value, n.value

    def accum_tensor(self):
        """The per-GPU accumulation buffer (rgb sum + count per pixel) as a torch CUDA tensor view, for the
        one NCCL reduce of the multi-GPU path (b2rt/dist.py)."""
        import torch
        ptr, n = self.accum_device_ptr()

        class _Wrap:
            pass
        w = _Wrap()
        w.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f4", "data": (ptr, False), "version": 2}
        dev = self.cfg.device if self.cfg.device >= 0 else torch.cuda.current_device()
        return torch.as_tensor(w, device=f"cuda:{dev}")


class Comm:
    """NCCL communicator behind the C ABI (b2rt_comm_*): one rank per GPU.

    Comm.unique_id() on rank 0 -> 128 bytes, shipped to the other ranks by any means (bench.py broadcasts them with
    torch.distributed), then Comm(n_ranks, rank, id, device) on every rank."""

    @staticmethod
    def _prefer_framework_nccl():
        # libb2rt.so binds NCCL at run time and takes the copy the process already has.  PyTorch bundles its own
        # (newer) libnccl.so.2; if the system copy were loaded first, a later `import torch` would resolve against it and
        # fail.  So in a Python process the framework's copy goes first.
        try:
            import torch  # noqa: F401
        except ImportError:
            pass

    def __init__(self, n_ranks, rank, uid, device=-1):
        self._prefer_framework_nccl()
        if len(uid) != 128:
            raise ValueError("unique id must be 128 bytes")
        buf = (C.c_uint8 * 128).from_buffer_copy(bytes(uid))
        h = C.c_void_p()
        _check(lib().b2rt_comm_create(n_ranks, rank, buf, device, C.byref(h)))
        self._h = h
        self.n_ranks, self.rank = n_ranks, rank

    @staticmethod
    def unique_id():
        Comm._prefer_framework_nccl()
        buf = (C.c_uint8 * 128)()
        _check(lib().b2rt_comm_unique_id(buf))
        return bytes(buf)

    @staticmethod
    def version():
        Comm._prefer_framework_nccl()
        return lib().b2rt_comm_version()

    def close(self):
        if getattr(self, "_h", None):
            lib().b2rt_comm_destroy(self._h); self._h = None

    __del__ = close


class CudaRenderer:
    """Mirror of cutracer::CudaRenderer (src/cudaRenderer.h:173-272): progressive accumulation."""

    def __init__(self, samples_per_frame=2, max_ray_depth=3, ns_area_light=2, median_threshold=32, **kw):
        self.pt = PathTracer(ns_aa=samples_per_frame, max_ray_depth=max_ray_depth, ns_area_light=ns_area_light,
                             median_threshold=median_threshold, **kw)
        self.scene = None
        self.frames = 0

    def allocOutputImage(self, width, height):
        self.pt.set_frame_size(width, height)

    def loadScene(self, scene_or_path):
        self.scene = load_scene(scene_or_path) if isinstance(scene_or_path, str) else scene_or_path
        self.pt.set_scene(self.scene)

    def setup(self):
        if self.pt._camera is None:
            self.pt.set_camera(place_camera(self.scene, self.pt.width, self.pt.height))

    def setViewpoint(self, camera_or_origin, look_at=None):
        """setViewpoint(camera) or, like the reference (src/cudaRenderer.cu:1845-1870), setViewpoint(origin, lookAt)."""
        cam = camera_or_origin if look_at is None else camera_look_at(camera_or_origin, look_at)
        self.pt.set_camera(cam)          # resets accumulation, cudaRenderer.cu:1866-1869
        self.frames = 0

    def clearImage(self):
        self.pt.clear(); self.frames = 0

    def render(self):
        # every frame draws fresh samples: shift the sample window
        self.pt.set_config(sample_first=self.frames * self.pt.cfg.ns_aa)
        self.pt.render()
        self.frames += 1

    def getImage(self):
        return self.pt.image()   # renderer-owned, valid until the next call (src/cudaRenderer.h:196)
