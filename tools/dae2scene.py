#!/usr/bin/env python3
"""COLLADA (.dae, CMU462 profile) -> flat binary scene (.b2s) converter.

Host-side tool, run in the build container where the reference media is mounted; the GPU box
only ever sees the .b2s files under scenes/.  It restates what the reference's loader does:

  * Collada::ColladaParser::load / parse_node / parse_polymesh / parse_material / parse_light /
    parse_sphere / parse_camera        (src/collada/collada.cpp:117-951)
      - up_axis fix-up matrix           (:151-190)
      - node <matrix> only ('break' after the first matrix, :232-256)
      - material: CMU462 <extra> technique wins, else phong/diffuse (:864-951)
  * Application::load: camera direction quirk c_dir = unit(M * (view_dir, 1)) (src/application.cpp:366-367),
    sphere centre/scale (:474-479)
  * DynamicScene::Mesh: vertices transformed by the node matrix (src/dynamic_scene/mesh.cpp:21-46)
  * StaticScene::Mesh: only the first three vertices of a polygon are used because
    HalfedgeMesh::triangulate is a stub (src/static_scene/object.cpp:17-72, src/meshEdit.cpp:360-364)
  * Vertex::normal(): area-weighted face-normal sum, normalised (src/halfEdgeMesh.h:619-644)
  * DynamicScene::AreaLight: position/direction/dim_x/dim_y from the node matrix
    (src/dynamic_scene/area_light.h:12-24)

File layout (.b2s, little endian):
  "B2S1" u32 version(1) u32 n_tris u32 n_spheres u32 n_materials u32 n_lights
  f32 cam_dir[3] f32 hfov f32 vfov f32 bbox[6]
  f32 tri_verts[n_tris*9] f32 tri_normals[n_tris*9] u32 tri_material[n_tris]
  f32 spheres[n_spheres*4] u32 sphere_material[n_spheres]
  materials[n_materials] (i32 kind, f32 albedo[3], transmittance[3], emission[3], ior, roughness)
  lights[n_lights]       (i32 kind, f32 radiance[3], position[3], direction[3], dim_x[3], dim_y[3])
"""
import math
import struct
import sys
import xml.etree.ElementTree as ET

import numpy as np

NS = "{http://www.collada.org/2005/11/COLLADASchema}"

MAT_DIFFUSE, MAT_MIRROR, MAT_GLASS, MAT_EMISSION, MAT_REFRACTION, MAT_GLOSSY = 0, 1, 2, 3, 4, 5
LIGHT_AREA, LIGHT_POINT, LIGHT_DIRECTIONAL = 0, 1, 2


def _find(e, path):
    return e.find("/".join(NS + p for p in path.split("/")))


def _floats(text):
    return np.array(text.split(), dtype=np.float64)


def _technique(e, profile):
    """get_technique_common / get_technique_cmu462: search the subtree for the technique."""
    if profile == "common":
        for t in e.iter(NS + "technique_common"):
            return t
        for t in e.iter(NS + "profile_COMMON"):
            tt = t.find(NS + "technique")
            if tt is not None:
                return tt
        return None
    for t in e.iter(NS + "technique"):
        if t.get("profile") == "CMU462":
            return t
    return None


class Scene:
    def __init__(self):
        self.tri_verts = []
        self.tri_normals = []
        self.tri_material = []
        self.spheres = []
        self.sphere_material = []
        self.materials = []
        self.lights = []
        self.cam_dir = np.array([0.0, 0.0, 1.0])
        self.hfov = 50.0
        self.vfov = 35.0
        self.bbox_min = np.full(3, np.inf)
        self.bbox_max = np.full(3, -np.inf)


def _material(root_ids, mat_id):
    m = root_ids[mat_id]
    inst = m.find(NS + "instance_effect")
    eff = root_ids[inst.get("url")[1:]]
    out = dict(kind=MAT_DIFFUSE, albedo=[0.5, 0.5, 0.5], transmittance=[0, 0, 0], emission=[0, 0, 0],
               ior=1.0, roughness=0.0)
    t462 = _technique(eff, "CMU462")
    tcom = _technique(eff, "common")
    if t462 is not None:
        for b in list(t462):
            tag = b.tag.replace(NS, "")
            get = lambda name: _floats(b.find(NS + name).text)
            if tag == "emission":
                rad = get("radiance")[:3]
                out.update(kind=MAT_EMISSION, emission=list(rad), albedo=[0, 0, 0])
            elif tag == "mirror":
                out.update(kind=MAT_MIRROR, albedo=list(get("reflectance")[:3]))
            elif tag == "glossy":       # commented out in the reference's parser (collada.cpp:898-907); accepted here
                out.update(kind=MAT_GLOSSY, albedo=list(get("reflectance")[:3]), roughness=float(get("roughness")[0]))
            elif tag == "refraction":
                out.update(kind=MAT_REFRACTION, transmittance=list(get("transmittance")[:3]),
                           roughness=float(get("roughness")[0]), ior=float(get("ior")[0]), albedo=[0, 0, 0])
            elif tag == "glass":
                out.update(kind=MAT_GLASS, transmittance=list(get("transmittance")[:3]),
                           albedo=list(get("reflectance")[:3]), roughness=float(get("roughness")[0]),
                           ior=float(get("ior")[0]))
    elif tcom is not None:
        d = _find(tcom, "phong/diffuse/color")
        if d is not None:
            out.update(albedo=list(_floats(d.text)[:3]))
    return out


def load_dae(path):
    tree = ET.parse(path)
    root = tree.getroot()
    ids = {e.get("id"): e for e in root.iter() if e.get("id") is not None}
    sc = Scene()

    up_axis = _find(root, "asset/up_axis").text.strip()
    G = np.eye(4)
    if up_axis == "X_UP":
        G[0, 0] = 0; G[0, 1] = 1; G[1, 0] = 1; G[1, 1] = 0; G[2, 2] = -1
    elif up_axis == "Z_UP":
        G[1, 1] = 0; G[1, 2] = 1; G[2, 1] = 1; G[2, 2] = 0; G[0, 0] = -1

    vs_inst = _find(root, "scene/instance_visual_scene")
    vscene = ids[vs_inst.get("url")[1:]]
    mat_index = {}

    def mat_for(node):
        im = _find(node, "instance_geometry/bind_material/technique_common/instance_material")
        if im is None:
            key = "__default_white__"
            if key not in mat_index:
                mat_index[key] = len(sc.materials)
                sc.materials.append(dict(kind=MAT_DIFFUSE, albedo=[1, 1, 1], transmittance=[0, 0, 0],
                                         emission=[0, 0, 0], ior=1.0, roughness=0.0))
            return mat_index[key]
        key = im.get("target")[1:]
        if key not in mat_index:
            mat_index[key] = len(sc.materials)
            sc.materials.append(_material(ids, key))
        return mat_index[key]

    def xf_point(M, p):
        v = M @ np.array([p[0], p[1], p[2], 1.0])
        return v[:3]

    def parse_node(node, parent):
        M = np.eye(4)
        for e in list(node):
            tag = e.tag.replace(NS, "")
            if tag == "matrix":
                vals = _floats(e.text)
                if len(vals) != 16:  # CBgems.dae ships a 15-entry camera matrix; pad from identity
                    pad = np.eye(4).reshape(-1)
                    pad[:min(len(vals), 16)] = vals[:16]
                    vals = pad
                M = vals.reshape(4, 4)
                break
            if tag == "translate":
                T = np.eye(4); T[:3, 3] = _floats(e.text)[:3]; M = T @ M
            # rotate/scale lists: the reference parser mis-reads these (collada.cpp:259-321);
            # none of the bundled path-tracer scenes use them.
        M = parent @ M
        for ch in node.findall(NS + "node"):
            parse_node(ch, M)
        icam = node.find(NS + "instance_camera")
        ilight = node.find(NS + "instance_light")
        igeom = node.find(NS + "instance_geometry")
        if icam is not None:
            cam = ids[icam.get("url")[1:]]
            persp = _find(cam, "optics/technique_common/perspective")
            xf = persp.find(NS + "xfov"); yf = persp.find(NS + "yfov")
            sc.hfov = float(xf.text) if xf is not None else 50.0
            sc.vfov = float(yf.text) if yf is not None else 35.0
            if yf is None:
                ar = float(persp.find(NS + "aspect_ratio").text)
                sc.vfov = 2 * math.degrees(math.atan(math.tan(math.radians(0.5 * sc.hfov)) / ar))
            d = xf_point(M, (0, 0, -1))
            sc.cam_dir = d / np.linalg.norm(d)
        elif ilight is not None:
            light = ids[ilight.get("url")[1:]]
            t = _technique(light, "CMU462")
            if t is None:
                t = _technique(light, "common")
            first = list(t)[0]
            kind = first.tag.replace(NS, "")
            col = _floats(first.find(NS + "color").text)[:3]
            pos = xf_point(M, (0, 0, 0))
            dirn = xf_point(M, (0, 0, -1)) - pos
            dirn = dirn / np.linalg.norm(dirn)
            up = np.array([0.0, 1.0, 0.0]); ldir = np.array([0.0, 0.0, -1.0])
            dim_y = xf_point(M, up) - pos
            dim_x = xf_point(M, np.cross(up, ldir)) - pos
            k = {"area": LIGHT_AREA, "point": LIGHT_POINT, "directional": LIGHT_DIRECTIONAL}.get(kind)
            if k is None:
                print(f"  (skipping unsupported light type '{kind}')", file=sys.stderr)
            else:
                sc.lights.append(dict(kind=k, radiance=list(col), position=list(pos), direction=list(dirn),
                                      dim_x=list(dim_x), dim_y=list(dim_y)))
        elif igeom is not None:
            geom = ids[igeom.get("url")[1:]]
            mesh = geom.find(NS + "mesh")
            if mesh is not None:
                add_mesh(mesh, M, mat_for(node))
            elif geom.find(NS + "extra") is not None:
                t = _technique(geom, "CMU462")
                r = float(_find(t, "sphere/radius").text)
                c = xf_point(M, (0, 0, 0))
                s = np.linalg.norm((M @ np.array([1.0, 0, 0, 0]))[:3])
                sc.spheres.append([c[0], c[1], c[2], r * s])
                sc.sphere_material.append(mat_for(node))
                sc.bbox_min = np.minimum(sc.bbox_min, c - r * s)
                sc.bbox_max = np.maximum(sc.bbox_max, c + r * s)

    def add_mesh(mesh, M, mat):
        sources = {}
        for s in mesh.findall(NS + "source"):
            fa = s.find(NS + "float_array")
            if fa is not None:
                sources[s.get("id")] = _floats(fa.text)
        vertices_e = mesh.find(NS + "vertices")
        pos_src = None
        for inp in vertices_e.findall(NS + "input"):
            if inp.get("semantic") == "POSITION":
                pos_src = inp.get("source")[1:]
        P = sources[pos_src].reshape(-1, 3)
        P = (np.c_[P, np.ones(len(P))] @ M.T)
        P = P[:, :3] / P[:, 3:4]
        pl = mesh.find(NS + "polylist")
        is_poly = pl is not None
        if pl is None:
            pl = mesh.find(NS + "triangles")
        has = {"VERTEX": None, "NORMAL": None, "TEXCOORD": None}
        for inp in pl.findall(NS + "input"):
            if inp.get("semantic") in has:
                has[inp.get("semantic")] = int(inp.get("offset"))
        stride = sum(1 for v in has.values() if v is not None)
        npoly = int(pl.get("count"))
        if is_poly:
            sizes = np.array(pl.find(NS + "vcount").text.split(), dtype=np.int64)[:npoly]
        else:
            sizes = np.full(npoly, 3, dtype=np.int64)
        idx = np.array(pl.find(NS + "p").text.split(), dtype=np.int64)
        starts = np.concatenate([[0], np.cumsum(sizes)[:-1]])
        voff = has["VERTEX"]
        # first three vertices of every polygon (triangulate() is a stub in the reference)
        tri = np.stack([idx[(starts + k) * stride + voff] for k in range(3)], axis=1)
        a, b, c = P[tri[:, 0]], P[tri[:, 1]], P[tri[:, 2]]
        # Vertex::normal(): sum over incident faces of cross(pj - pi, pk - pi) == the face's
        # (un-normalised) normal for every corner of a triangle
        fn = np.cross(b - a, c - a)
        N = np.zeros_like(P)
        for k in range(3):
            np.add.at(N, tri[:, k], fn)
        ln = np.linalg.norm(N, axis=1, keepdims=True)
        ln[ln == 0] = 1.0
        N = N / ln
        sc.tri_verts.append(np.stack([a, b, c], axis=1).reshape(-1, 9))
        sc.tri_normals.append(np.stack([N[tri[:, 0]], N[tri[:, 1]], N[tri[:, 2]]], axis=1).reshape(-1, 9))
        sc.tri_material.append(np.full(len(tri), mat, dtype=np.uint32))
        sc.bbox_min = np.minimum(sc.bbox_min, P.min(axis=0))
        sc.bbox_max = np.maximum(sc.bbox_max, P.max(axis=0))

    for n in vscene.findall(NS + "node"):
        parse_node(n, G)
    return sc


def pack_scene(sc):
    tv = np.concatenate(sc.tri_verts).astype(np.float32) if sc.tri_verts else np.zeros((0, 9), np.float32)
    tn = np.concatenate(sc.tri_normals).astype(np.float32) if sc.tri_normals else np.zeros((0, 9), np.float32)
    tm = np.concatenate(sc.tri_material).astype(np.uint32) if sc.tri_material else np.zeros(0, np.uint32)
    sp = np.array(sc.spheres, dtype=np.float32).reshape(-1, 4)
    sm = np.array(sc.sphere_material, dtype=np.uint32)
    out = bytearray()
    out += b"B2S1" + struct.pack("<5I", 1, len(tv), len(sp), len(sc.materials), len(sc.lights))
    out += struct.pack("<3f2f6f", *sc.cam_dir, sc.hfov, sc.vfov, *sc.bbox_min, *sc.bbox_max)
    out += tv.tobytes() + tn.tobytes() + tm.tobytes() + sp.tobytes() + sm.tobytes()
    for m in sc.materials:
        out += struct.pack("<i11f", m["kind"], *m["albedo"], *m["transmittance"], *m["emission"], m["ior"],
                           m["roughness"])
    for l in sc.lights:
        out += struct.pack("<i15f", l["kind"], *l["radiance"], *l["position"], *l["direction"], *l["dim_x"],
                           *l["dim_y"])
    return bytes(out)


def main():
    if len(sys.argv) != 3:
        print("usage: dae2scene.py in.dae out.b2s")
        return 2
    sc = load_dae(sys.argv[1])
    data = pack_scene(sc)
    with open(sys.argv[2], "wb") as f:
        f.write(data)
    nt = sum(len(t) for t in sc.tri_verts)
    print(f"{sys.argv[1]}: {nt} tris, {len(sc.spheres)} spheres, {len(sc.materials)} materials, "
          f"{len(sc.lights)} lights, bbox {sc.bbox_min} .. {sc.bbox_max}, cam_dir {sc.cam_dir} -> {sys.argv[2]}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
