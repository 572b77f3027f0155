"""Flat scene container (.b2s files, synthetic generators) and camera placement (host logic, no GPU).

Camera placement restates Application::load (src/application.cpp:395-408) + Camera::configure /
place / compute_position (src/camera.cpp:15-33,35-50,87-109) of the reference.
"""
import ctypes as C
import math
import struct

import numpy as np

from ._abi import Camera, Light, Material, SceneDesc

MAT_DIFFUSE, MAT_MIRROR, MAT_GLASS, MAT_EMISSION, MAT_REFRACTION, MAT_GLOSSY = 0, 1, 2, 3, 4, 5
LIGHT_AREA, LIGHT_POINT, LIGHT_DIRECTIONAL = 0, 1, 2


class Scene:
    """Host copy of a scene in the b2rt_scene_desc layout (include/b2rt.h)."""

    def __init__(self, tri_verts, tri_normals=None, tri_material=None, spheres=None, sphere_material=None,
                 materials=None, lights=None, cam_dir=(0, 0, 1), hfov=50.0, vfov=35.0, bbox=None):
        self.tri_verts = np.ascontiguousarray(tri_verts, dtype=np.float32).reshape(-1, 9)
        n = len(self.tri_verts)
        self.tri_normals = None if tri_normals is None else np.ascontiguousarray(tri_normals, np.float32).reshape(n, 9)
        self.tri_material = (np.zeros(n, np.uint32) if tri_material is None
                             else np.ascontiguousarray(tri_material, np.uint32))
        self.spheres = (np.zeros((0, 4), np.float32) if spheres is None
                        else np.ascontiguousarray(spheres, np.float32).reshape(-1, 4))
        self.sphere_material = (np.zeros(len(self.spheres), np.uint32) if sphere_material is None
                                else np.ascontiguousarray(sphere_material, np.uint32))
        self.materials = materials or [dict(kind=MAT_DIFFUSE, albedo=(0.5, 0.5, 0.5))]
        self.lights = lights or []
        self.cam_dir = np.asarray(cam_dir, np.float64)
        self.hfov, self.vfov = float(hfov), float(vfov)
        if bbox is None:
            lo = np.full(3, np.inf); hi = np.full(3, -np.inf)
            if n:
                p = self.tri_verts.reshape(-1, 3)
                lo = np.minimum(lo, p.min(0)); hi = np.maximum(hi, p.max(0))
            for s in self.spheres:
                lo = np.minimum(lo, s[:3] - s[3]); hi = np.maximum(hi, s[:3] + s[3])
            bbox = np.concatenate([lo, hi])
        self.bbox = np.asarray(bbox, np.float64)

    @property
    def n_tris(self):
        return len(self.tri_verts)

    @property
    def n_prims(self):
        return len(self.tri_verts) + len(self.spheres)

    def desc(self):
        """Returns (SceneDesc, keepalive) -- keepalive must outlive the call that consumes the desc."""
        mats = (Material * len(self.materials))()
        for i, m in enumerate(self.materials):
            mats[i].kind = int(m.get("kind", 0))
            mats[i].albedo[:] = [float(x) for x in m.get("albedo", (0, 0, 0))]
            mats[i].transmittance[:] = [float(x) for x in m.get("transmittance", (0, 0, 0))]
            mats[i].emission[:] = [float(x) for x in m.get("emission", (0, 0, 0))]
            mats[i].ior = float(m.get("ior", 1.0)); mats[i].roughness = float(m.get("roughness", 0.0))
        lights = (Light * max(1, len(self.lights)))()
        for i, l in enumerate(self.lights):
            lights[i].kind = int(l.get("kind", 0))
            for k in ("radiance", "position", "direction", "dim_x", "dim_y"):
                getattr(lights[i], k)[:] = [float(x) for x in l.get(k, (0, 0, 0))]
        d = SceneDesc()
        fp = C.POINTER(C.c_float); up = C.POINTER(C.c_uint32)
        d.n_tris = self.n_tris
        d.tri_verts = self.tri_verts.ctypes.data_as(fp)
        d.tri_normals = self.tri_normals.ctypes.data_as(fp) if self.tri_normals is not None else None
        d.tri_material = self.tri_material.ctypes.data_as(up)
        d.n_spheres = len(self.spheres)
        d.spheres = self.spheres.ctypes.data_as(fp)
        d.sphere_material = self.sphere_material.ctypes.data_as(up)
        d.n_materials = len(self.materials); d.materials = mats
        d.n_lights = len(self.lights); d.lights = lights
        return d, (mats, lights, self)

    # ---- .b2s files (layout documented in tools/dae2scene.py) ----
    @staticmethod
    def load(path):
        with open(path, "rb") as f:
            data = f.read()
        if data[:4] != b"B2S1":
            raise ValueError(f"{path}: not a .b2s scene")
        ver, nt, ns, nm, nl = struct.unpack_from("<5I", data, 4)
        off = 24
        hdr = struct.unpack_from("<11f", data, off); off += 44
        def arr(dt, count):
            nonlocal off
            a = np.frombuffer(data, dtype=dt, count=count, offset=off).copy()
            off += a.nbytes
            return a
        tv = arr(np.float32, nt * 9); tn = arr(np.float32, nt * 9); tm = arr(np.uint32, nt)
        sp = arr(np.float32, ns * 4); sm = arr(np.uint32, ns)
        mats = []
        for _ in range(nm):
            v = struct.unpack_from("<i11f", data, off); off += 48
            mats.append(dict(kind=v[0], albedo=v[1:4], transmittance=v[4:7], emission=v[7:10], ior=v[10], roughness=v[11]))
        lights = []
        for _ in range(nl):
            v = struct.unpack_from("<i15f", data, off); off += 64
            lights.append(dict(kind=v[0], radiance=v[1:4], position=v[4:7], direction=v[7:10], dim_x=v[10:13],
                               dim_y=v[13:16]))
        return Scene(tv, tn, tm, sp, sm, mats, lights, cam_dir=hdr[0:3], hfov=hdr[3], vfov=hdr[4], bbox=hdr[5:11])


def place_camera(scene, width, height):
    """Application::load camera placement for a frame of width x height -> _abi.Camera."""
    bbox = scene.bbox
    lo, hi = bbox[:3], bbox[3:]
    target = (lo + hi) / 2
    canonical = np.linalg.norm(hi - lo) / 2 * 1.5            # application.cpp:398
    r = canonical * 2
    r = min(max(r, canonical / 10.0), canonical * 20.0)      # camera.cpp:38
    c_dir = scene.cam_dir / np.linalg.norm(scene.cam_dir)
    phi = math.acos(max(-1.0, min(1.0, c_dir[1])))
    theta = math.atan2(c_dir[0], c_dir[2])
    if math.sin(phi) == 0:
        phi += 1e-5                                           # EPS_F
    sp = math.sin(phi)
    to_cam = np.array([r * sp * math.sin(theta), r * math.cos(phi), r * sp * math.cos(theta)])
    pos = target + to_cam
    up = np.array([0.0, 1.0 if sp > 0 else -1.0, 0.0])
    xdir = np.cross(up, to_cam); xdir /= np.linalg.norm(xdir)
    ydir = np.cross(to_cam, xdir); ydir /= np.linalg.norm(ydir)
    zdir = to_cam / np.linalg.norm(to_cam)
    # Camera::configure fov fix-up, camera.cpp:15-33
    hfov, vfov = scene.hfov, scene.vfov
    ar1 = math.tan(math.radians(hfov) / 2) / math.tan(math.radians(vfov) / 2)
    ar = width / height
    if ar1 < ar:
        hfov = 2 * math.degrees(math.atan(math.tan(math.radians(vfov) / 2) * ar))
    elif ar1 > ar:
        vfov = 2 * math.degrees(math.atan(math.tan(math.radians(hfov) / 2) / ar))
    cam = Camera()
    cam.pos[:] = [float(x) for x in pos]
    cam.c2w[:] = [float(x) for x in np.concatenate([xdir, ydir, zdir])]
    cam.hfov_deg = hfov; cam.vfov_deg = vfov
    return cam


# ---- synthetic scenes (deterministic; SURVEY 8d cfg3/cfg5 stand-ins) ----
def random_soup(n_tris, size=0.01, seed=0x5EED0001):
    """cfg5: triangle soup in the unit cube: vertices c + 0.5*s*u_k, c ~ U[0,1)^3, u_k ~ U(-1,1)^3."""
    rng = np.random.Generator(np.random.Philox(seed))
    c = rng.random((n_tris, 1, 3), dtype=np.float32)
    u = rng.random((n_tris, 3, 3), dtype=np.float32) * 2 - 1
    v = (c + np.float32(0.5 * size) * u).astype(np.float32)
    return Scene(v.reshape(n_tris, 9), cam_dir=(0, 0, 1))


def subdivide(scene, levels=1, select=None):
    """Midpoint-subdivide the selected triangles (1 -> 4), keeping materials; normals are re-blended.
    Used for the 'dragon-class' stand-in: CBbunny with the bunny subdivided once."""
    tv, tn, tm = scene.tri_verts.reshape(-1, 3, 3), scene.tri_normals, scene.tri_material
    tn = None if tn is None else tn.reshape(-1, 3, 3)
    for _ in range(levels):
        sel = np.ones(len(tv), bool) if select is None else select(tv, tm)
        def split(a):
            p0, p1, p2 = a[:, 0], a[:, 1], a[:, 2]
            m01, m12, m20 = (p0 + p1) * 0.5, (p1 + p2) * 0.5, (p2 + p0) * 0.5
            return np.stack([np.stack([p0, m01, m20], 1), np.stack([m01, p1, m12], 1),
                             np.stack([m20, m12, p2], 1), np.stack([m01, m12, m20], 1)], 1).reshape(-1, 3, 3)
        new_v = np.concatenate([tv[~sel], split(tv[sel])])
        new_m = np.concatenate([tm[~sel], np.repeat(tm[sel], 4)])
        if tn is not None:
            nn = split(tn[sel])
            nn /= np.maximum(np.linalg.norm(nn, axis=2, keepdims=True), 1e-20)
            tn = np.concatenate([tn[~sel], nn]).astype(np.float32)
        tv, tm = new_v.astype(np.float32), new_m
    return Scene(tv.reshape(-1, 9), None if tn is None else tn.reshape(-1, 9), tm, scene.spheres,
                 scene.sphere_material, scene.materials, scene.lights, scene.cam_dir, scene.hfov, scene.vfov,
                 scene.bbox)


def camera_rays(cam, width, height, jitter=None):
    """Pixel-centre camera rays (generate_ray contract, src/camera.h:71-81) as (org[n,3], dir[n,3]) fp32."""
    xs = (np.arange(width, dtype=np.float32) + np.float32(0.5)) / np.float32(width)
    ys = (np.arange(height, dtype=np.float32) + np.float32(0.5)) / np.float32(height)
    sx, sy = np.meshgrid(xs, ys)
    th = np.float32(math.tan(math.radians(cam.hfov_deg) / 2)); tv = np.float32(math.tan(math.radians(cam.vfov_deg) / 2))
    px = (2 * sx - 1) * th; py = (2 * sy - 1) * tv
    c2w = np.array(cam.c2w[:], np.float32).reshape(3, 3)
    d = px[..., None] * c2w[0] + py[..., None] * c2w[1] - c2w[2]
    d = d / np.linalg.norm(d, axis=-1, keepdims=True)
    o = np.broadcast_to(np.array(cam.pos[:], np.float32), d.shape)
    return np.ascontiguousarray(o.reshape(-1, 3), np.float32), np.ascontiguousarray(d.reshape(-1, 3), np.float32)


def cfg3_standin(cbbunny):
    """BASELINE configs[2] stand-in (the dragon asset is missing from the reference checkout, .MISSING_LARGE_BLOBS):
    CBbunny with the bunny mesh midpoint-subdivided once -> 114,316 triangles."""
    return subdivide(cbbunny, 1, select=lambda tv, tm: tm == np.argmax(np.bincount(tm)))


def cfg4_standin(cbbunny, levels=2):
    """BASELINE configs[3] stand-in ("Lucy / largest bundled mesh, mirror + glass BSDFs"; no such asset is bundled): the
    CBbunny Cornell box with the bunny subdivided twice (457,228 triangles) and made of the glass of CBspheres.dae
    (ior 1.45), plus two analytic mirror spheres of CBspheres.dae's mirror material beside it."""
    sc = subdivide(cbbunny, levels, select=lambda tv, tm: tm == np.argmax(np.bincount(tm))) if levels else cbbunny
    mats = [dict(m) for m in sc.materials]
    glass = int(np.argmax(np.bincount(sc.tri_material)))
    mats[glass] = dict(kind=MAT_GLASS, albedo=(1.0, 1.0, 1.0), transmittance=(1.0, 1.0, 1.0), ior=1.45)
    mats.append(dict(kind=MAT_MIRROR, albedo=(1.0, 1.0, 1.0)))
    spheres = np.array([[0.68, 0.25, 0.45, 0.25], [-0.66, 0.3, -0.45, 0.3]], np.float32)
    return Scene(sc.tri_verts, sc.tri_normals, sc.tri_material, spheres, np.full(2, len(mats) - 1, np.uint32), mats, sc.lights,
                 sc.cam_dir, sc.hfov, sc.vfov, sc.bbox)
