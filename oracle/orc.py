"""ctypes wrapper of oracle/liboracle.so -- TEST INFRASTRUCTURE (see oracle/oracle.cpp header).
Import only from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs."""
import ctypes as C
import importlib
import os
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(os.path.dirname(_HERE), "cuda-raytracer_b200"))
_abi = importlib.import_module("b2rt._abi")

_lib = None


def lib():
    global _lib
    if _lib is None:
        path = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(path):
            import subprocess
            subprocess.check_call(["make", "-C", _HERE, "-s"])
        _lib = C.CDLL(path)
        _lib.orc_scene_create.restype = C.c_void_p
        _lib.orc_scene_create.argtypes = [C.POINTER(_abi.SceneDesc), C.c_uint32]
        _lib.orc_scene_destroy.argtypes = [C.c_void_p]
        _lib.orc_bvh_node_count.argtypes = [C.c_void_p]; _lib.orc_bvh_node_count.restype = C.c_uint32
        _lib.orc_bvh_dump.argtypes = [C.c_void_p] + [C.c_void_p] * 5
        _lib.orc_wide_levels.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32]; _lib.orc_wide_levels.restype = C.c_uint32
        _lib.orc_intersect.argtypes = [C.c_void_p, C.c_int, C.c_int] + [C.c_void_p] * 4 + [C.c_uint64, C.c_void_p,
                                                                                           C.c_void_p, C.c_int, C.c_void_p]
        _lib.orc_render.argtypes = [C.c_void_p, C.POINTER(_abi.Camera), C.POINTER(_abi.Config), C.c_uint32, C.c_uint32,
                                    C.c_int, C.c_void_p, C.c_void_p, C.POINTER(C.c_double), C.c_uint32]
        _lib.orc_tonemap.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p]
        _lib.orc_median3x3.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p]
        _lib.orc_filter.argtypes = [C.c_uint32, C.c_float, C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p]
        _lib.orc_philox.argtypes = [C.c_uint32] * 6 + [C.c_void_p]
        _lib.orc_sincos2pi.argtypes = [C.c_float, C.POINTER(C.c_float), C.POINTER(C.c_float)]
        _lib.orc_set_envmap.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32]
        _lib.orc_atan2_turns.argtypes = [C.c_float, C.c_float]; _lib.orc_atan2_turns.restype = C.c_float
    return _lib


class OracleScene:
    def __init__(self, scene, max_leaf=4):
        """max_leaf=None: no BVH (the reference builder restated here is O(n log^2 n)); only mode="brute" queries."""
        d, keep = scene.desc()
        self._h = lib().orc_scene_create(C.byref(d), 0xFFFFFFFF if max_leaf is None else max_leaf)
        self.scene = scene
        self.n_prims = scene.n_prims

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_scene_destroy(self._h); self._h = None

    def bvh_dump(self):
        n = lib().orc_bvh_node_count(self._h)
        order = np.zeros(self.n_prims, np.uint32)
        start = np.zeros(n, np.uint64); rng = np.zeros(n, np.uint64)
        left = np.zeros(n, np.int32); right = np.zeros(n, np.int32)
        lib().orc_bvh_dump(self._h, order.ctypes.data, start.ctypes.data, rng.ctypes.data, left.ctypes.data,
                           right.ctypes.data)
        return dict(order=order, start=start, range=rng, left=left, right=right)

    def wide_levels(self):
        buf = np.zeros(64, np.uint32)
        n = lib().orc_wide_levels(self._h, buf.ctypes.data, 64)
        return buf[:n].tolist()

    def intersect(self, org, dirs, tmin=None, tmax=None, mode="bvh", any_hit=False, threads=None):
        org = np.ascontiguousarray(org, np.float32); dirs = np.ascontiguousarray(dirs, np.float32)
        n = len(org)
        tmin = np.zeros(n, np.float32) if tmin is None else np.ascontiguousarray(tmin, np.float32)
        tmax = np.full(n, np.inf, np.float32) if tmax is None else np.ascontiguousarray(tmax, np.float32)
        t = np.zeros(n, np.float32); prim = np.zeros(n, np.uint32); cnt = np.zeros(2, np.uint64)
        lib().orc_intersect(self._h, 0 if mode == "brute" else 1, int(any_hit), org.ctypes.data, dirs.ctypes.data,
                            tmin.ctypes.data, tmax.ctypes.data, n, t.ctypes.data, prim.ctypes.data,
                            threads or os.cpu_count() or 1, cnt.ctypes.data)
        self.last_counters = dict(box_tests=int(cnt[0]), prim_tests=int(cnt[1]))
        return t, prim

    def set_envmap(self, rgb):
        """float32 [h, w, 3] environment map for render() (None removes it); layout as b2rt_set_envmap."""
        if rgb is None:
            lib().orc_set_envmap(self._h, None, 0, 0)
        else:
            a = np.ascontiguousarray(rgb, np.float32)
            lib().orc_set_envmap(self._h, a.ctypes.data, a.shape[1], a.shape[0])

    def render(self, cam, cfg, width, height, threads=None, tile_stride=1):
        rgb = np.zeros((height, width, 3), np.float32)
        stats = np.zeros(5, np.uint64); sec = C.c_double(0)
        lib().orc_render(self._h, C.byref(cam), C.byref(cfg), width, height, threads or os.cpu_count() or 1,
                         rgb.ctypes.data, stats.ctypes.data, C.byref(sec), tile_stride)
        self.last_stats = dict(rays_camera=int(stats[0]), rays_bounce=int(stats[1]), rays_shadow=int(stats[2]),
                               box_tests=int(stats[3]), prim_tests=int(stats[4]), seconds=sec.value)
        return rgb


def tonemap(rgb):
    rgb = np.ascontiguousarray(rgb, np.float32)
    out = np.zeros(rgb.shape[:-1], np.uint32)
    lib().orc_tonemap(rgb.ctypes.data, out.size, out.ctypes.data)
    return out


def recon_filter(rgb, kind, sigma_r=0.0):
    """b2rt_config.filter_kind 1 (3x3 Gaussian) / 2 (5x5 joint bilateral)."""
    rgb = np.ascontiguousarray(rgb, np.float32)
    out = np.zeros_like(rgb)
    lib().orc_filter(kind, sigma_r, rgb.ctypes.data, rgb.shape[1], rgb.shape[0], out.ctypes.data)
    return out


def median3x3(rgb):
    rgb = np.ascontiguousarray(rgb, np.float32)
    out = np.zeros_like(rgb)
    lib().orc_median3x3(rgb.ctypes.data, rgb.shape[1], rgb.shape[0], out.ctypes.data)
    return out


def philox(c, k):
    out = np.zeros(4, np.uint32)
    lib().orc_philox(*[int(x) for x in c], int(k[0]), int(k[1]), out.ctypes.data)
    return out


def atan2_turns(y, x):
    return float(lib().orc_atan2_turns(float(y), float(x)))


def sincos2pi(u):
    s = C.c_float(); c = C.c_float()
    lib().orc_sincos2pi(float(u), C.byref(s), C.byref(c))
    return s.value, c.value
