"""CPU tests of the product's host side: the C-ABI library loads and exports every symbol of include/b2rt.h,
fails loudly without a device, the host BVH builder produces structurally valid subtree blobs, scene IO and
camera placement agree with the Python mirror, and the multi-GPU plumbing (sample sharding + one reduce)
works with world_size 2 over gloo."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest

import b2rt
from b2rt._abi import Camera, Config
from b2rt.scene import Scene, place_camera, random_soup, subdivide
from conftest import ROOT, scene_path


def test_header_symbols_exported():
    hdr = open(os.path.join(ROOT, "include", "b2rt.h")).read()
    declared = set(re.findall(r"\b(b2rt_[a-z0-9_]+)\s*\(", hdr))
    declared -= {"b2rt_status"}
    lib = b2rt.lib()
    missing = [s for s in sorted(declared) if not hasattr(lib, s)]
    assert not missing, missing
    assert set(b2rt.EXPORTS) <= declared
    assert lib.b2rt_abi_version() == 2


def test_fails_loudly_without_device():
    if b2rt.device_count() > 0:
        pytest.skip("a CUDA device is visible")
    sc = Scene.load(scene_path("trigs1"))
    with pytest.raises(b2rt.B2rtError) as e:
        b2rt.BVHAccel(sc)
    assert e.value.code == -2 and "no CPU fallback" in str(e.value)
    with pytest.raises(b2rt.B2rtError) as e:
        b2rt.PathTracer()
    assert e.value.code == -2
    with pytest.raises(b2rt.B2rtError) as e:
        b2rt.BVHAccel(sc, builder="gpu")      # b2rt_bvh_build_device: the device builder has no host fallback either
    assert e.value.code == -2 and "no CPU fallback" in str(e.value)


def test_new_bvh_entry_points_reject_null_handles():
    lib = b2rt.lib()
    assert lib.b2rt_bvh_set_slicing(None, C.c_float(1.0), C.c_float(4.0), 4) == -1
    assert lib.b2rt_bvh_validate(None, None, None) == -1
    assert lib.b2rt_bvh_build_device(None, 4, 4, 0, -1, None) == -1


def test_product_does_not_link_oracle():
    out = subprocess.check_output(["ldd", b2rt.LIB_PATH], text=True)
    assert "oracle" not in out
    for root, _, files in os.walk(os.path.join(ROOT, "cuda-raytracer_b200")):
        for f in files:
            if f.endswith((".cu", ".cuh", ".cpp", ".h", ".py")):
                src = open(os.path.join(root, f)).read()
                assert "liboracle" not in src and "import orc" not in src, f


@pytest.mark.parametrize("name", ["CBbunny", "CBcoil", "CBgems", "CBspheres_lambertian", "CBempty", "trigs1", "sphere_diffuse",
                                  "plane1024"])
@pytest.mark.parametrize("width", [2, 4, 8, 16])
def test_bvh_blob_structure(name, width):
    sc = Scene.load(scene_path(name))
    for tb in (0, 8192, 150000):
        for ml in (2, 4, 8):
            st = b2rt.validate_bvh_host(sc, ml, width, tb)
            assert st["stack_bound"] <= 64
            assert st["subtrees"] >= 1 and st["levels"] >= 1


def test_bvh_blob_structure_soup_and_degenerate():
    st = b2rt.validate_bvh_host(random_soup(200000), 4, 4, 0)
    assert st["levels"] >= 2 and st["exits"] == st["subtrees"] - 1
    # all-coincident triangles (degenerate centroids) and a single triangle
    tri = np.tile(np.array([[0, 0, 0, 1, 0, 0, 0, 1, 0]], np.float32), (1000, 1))
    st = b2rt.validate_bvh_host(Scene(tri), 4, 4, 0)
    assert st["leaves"] >= 250
    st = b2rt.validate_bvh_host(Scene(tri[:1]), 4, 8, 0)
    assert st["wide_nodes"] == 1 and st["leaves"] == 1


def test_invalid_arguments():
    sc = Scene.load(scene_path("trigs1"))
    with pytest.raises(b2rt.B2rtError):
        b2rt.validate_bvh_host(sc, 4, 5, 0)          # width must be 2, 4, 8 or 16
    with pytest.raises(b2rt.B2rtError):
        b2rt.validate_bvh_host(sc, 100, 4, 0)        # leaf size limit
    bad = Scene(sc.tri_verts, None, np.array([7], np.uint32))
    with pytest.raises(b2rt.B2rtError):
        b2rt.validate_bvh_host(bad, 4, 4, 0)         # material index out of range


def test_scene_file_roundtrip_and_camera(tmp_path):
    lib = b2rt.lib()
    SceneFile = b2rt._abi.SceneFile
    for name in ("CBspheres_lambertian", "CBcoil"):
        sc = Scene.load(scene_path(name))
        p = C.c_void_p()
        assert lib.b2rt_scene_load(scene_path(name).encode(), C.byref(p)) == 0
        f = C.cast(p, C.POINTER(SceneFile)).contents
        assert f.desc.n_tris == sc.n_tris and f.desc.n_spheres == len(sc.spheres) and f.desc.n_lights == len(sc.lights)
        tv = np.ctypeslib.as_array(f.desc.tri_verts, shape=(sc.n_tris * 9,))
        assert np.array_equal(tv, sc.tri_verts.reshape(-1))
        out = tmp_path / (name + ".b2s")
        assert lib.b2rt_scene_save(str(out).encode(), p) == 0
        assert open(out, "rb").read() == open(scene_path(name), "rb").read()
        for (w, h) in ((640, 480), (1920, 1080), (100, 300)):
            cam_c = Camera()
            bbox = (C.c_float * 6)(*[float(x) for x in sc.bbox]); vd = (C.c_float * 3)(*[float(x) for x in sc.cam_dir])
            assert lib.b2rt_camera_place(bbox, vd, sc.hfov, sc.vfov, w, h, C.byref(cam_c)) == 0
            cam_p = place_camera(sc, w, h)
            np.testing.assert_allclose(cam_c.pos[:], cam_p.pos[:], rtol=1e-6, atol=1e-6)
            np.testing.assert_allclose(cam_c.c2w[:], cam_p.c2w[:], rtol=1e-6, atol=1e-6)
            assert abs(cam_c.hfov_deg - cam_p.hfov_deg) < 1e-4 and abs(cam_c.vfov_deg - cam_p.vfov_deg) < 1e-4
        lib.b2rt_scene_free(p)
    p = C.c_void_p()
    assert lib.b2rt_scene_load(b"/nonexistent.b2s", C.byref(p)) == -5


def test_subdivide_standin():
    from b2rt.scene import cfg3_standin, cfg4_standin
    sc = Scene.load(scene_path("CBbunny"))
    big = cfg3_standin(sc)
    assert big.n_tris == 12 + 28576 * 4     # "CBdragon_standin" = 114,316 triangles
    c4 = cfg4_standin(sc)                    # BASELINE configs[3] stand-in: glass mesh subdivided twice + mirror spheres
    assert c4.n_tris == 12 + 28576 * 16 and len(c4.spheres) == 2
    kinds = [m["kind"] for m in c4.materials]
    assert 2 in kinds and 1 in kinds         # MAT_GLASS, MAT_MIRROR
    st = b2rt.validate_bvh_host(big, 4, 4, 0)
    assert st["leaves"] > 20000


def test_shard_samples():
    from b2rt.dist import shard_samples
    for total in (1, 7, 64, 256):
        for world in (1, 2, 3, 8):
            got = []
            for r in range(world):
                first, stride, cnt = shard_samples(total, r, world)
                got += [first + k * stride for k in range(cnt)]
            assert sorted(got) == list(range(total))


_WORKER = r'''
import os, sys
sys.path.insert(0, os.path.join(sys.argv[1], "cuda-raytracer_b200")); sys.path.insert(0, os.path.join(sys.argv[1], "oracle"))
import numpy as np, torch, torch.distributed as dist
from b2rt._abi import Config
from b2rt.scene import Scene, place_camera
from b2rt.dist import shard_samples, reduce_accum, resolve_mean
import orc   # the CPU oracle stands in for the device renderer in this gloo test of the host plumbing
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
sc = Scene.load(os.path.join(sys.argv[1], "scenes", "CBspheres_lambertian.b2s"))
w, h, spp = 32, 24, 6
cam = place_camera(sc, w, h)
first, stride, cnt = shard_samples(spp, rank, world)
o = orc.OracleScene(sc, 4)
img = o.render(cam, Config(ns_aa=cnt, max_ray_depth=3, ns_area_light=1, seed=9, sample_first=first, sample_stride=stride), w, h, threads=1)
accum = torch.zeros(h * w * 4)
a = accum.view(-1, 4); a[:, :3] = torch.from_numpy(img.reshape(-1, 3)) * cnt; a[:, 3] = cnt
reduce_accum(accum, dst=0)
if rank == 0:
    full = o.render(cam, Config(ns_aa=spp, max_ray_depth=3, ns_area_light=1, seed=9), w, h, threads=1)
    got = resolve_mean(accum).numpy().reshape(h, w, 3)
    assert float(accum.view(-1, 4)[:, 3].min()) == spp
    np.testing.assert_allclose(got, full, rtol=2e-5, atol=1e-6)
    print("DIST_OK")
dist.destroy_process_group()
'''


def test_two_rank_gloo_reduce(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                          "--master-port", "29541", str(script), ROOT], capture_output=True, text=True, env=env, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    assert "DIST_OK" in out.stdout


def test_cpp_example_fails_loudly_without_device():
    exe = os.path.join(ROOT, "cuda-raytracer_b200", "examples", "render_scene")
    if not os.path.exists(exe):
        pytest.skip("C++ example not built")
    if b2rt.device_count() > 0:
        pytest.skip("a CUDA device is visible")
    out = subprocess.run([exe, scene_path("CBempty")], capture_output=True, text=True)
    assert out.returncode == 1 and "no CPU fallback" in out.stderr


# ---- b2rt_load_dae: the C++ COLLADA loader ---------------------------------------------
def _same_scene(a, b):
    assert np.array_equal(a.tri_verts, b.tri_verts)
    if a.n_tris:
        assert np.array_equal(a.tri_normals, b.tri_normals)
    assert np.array_equal(a.tri_material, b.tri_material)
    assert np.array_equal(a.spheres, b.spheres) and np.array_equal(a.sphere_material, b.sphere_material)
    assert a.materials == b.materials and a.lights == b.lights
    assert np.array_equal(a.cam_dir.astype(np.float32), b.cam_dir.astype(np.float32))
    assert a.hfov == b.hfov and a.vfov == b.vfov and np.array_equal(a.bbox, b.bbox)


def test_load_dae_fixture_matches_converter():
    """tests/golden/mini_scene.dae (hand-written) through the C++ loader == the .b2s the independent Python
    converter (tools/dae2scene.py) produced from the same file, bit for bit; also checks the loader's semantics."""
    a = b2rt.load_dae(os.path.join(ROOT, "tests", "golden", "mini_scene.dae"))
    b = Scene.load(os.path.join(ROOT, "tests", "golden", "mini_scene.b2s"))
    _same_scene(a, b)
    assert a.n_tris == 7 and len(a.spheres) == 1
    # Z_UP -> Y_UP fix-up (collada.cpp:171-183): the floor's z = 0.125 corner became y = 0.125; quad -> first 3 vertices
    assert a.tri_verts[2].tolist() == [2, 0, -2, -2, 0, -2, -2, 0.125, 2]
    kinds = [m["kind"] for m in a.materials]
    assert kinds == [3, 0, 1, 2, 0]            # emission, phong diffuse, mirror, glass, default white
    assert a.materials[4]["albedo"] == (1.0, 1.0, 1.0)
    assert [l["kind"] for l in a.lights] == [0, 2]     # CMU462 <area> wins over the common <point>; directional
    assert a.lights[0]["position"] == (-0.125, 2.0, -0.25)      # parent <translate> applied
    assert a.spheres[0].tolist() == [-0.25, 0.25, -0.5, 0.25]    # radius scaled by |M e_x|
    assert abs(a.vfov - 29.068184) < 1e-5                       # derived from xfov + aspect_ratio


def test_load_dae_roundtrip_save(tmp_path):
    lib = b2rt.lib()
    pf = C.c_void_p()
    assert lib.b2rt_load_dae(os.path.join(ROOT, "tests", "golden", "mini_scene.dae").encode(), C.byref(pf)) == 0
    out = str(tmp_path / "mini.b2s")
    assert lib.b2rt_scene_save(out.encode(), pf) == 0
    lib.b2rt_scene_free(pf)
    assert open(out, "rb").read() == open(os.path.join(ROOT, "tests", "golden", "mini_scene.b2s"), "rb").read()
    _same_scene(b2rt.load_scene(out), b2rt.load_scene(os.path.join(ROOT, "tests", "golden", "mini_scene.dae")))


@pytest.mark.parametrize("text,needle", [
    ("", "no root element"),
    ("<COLLADA><asset><up_axis>Y_UP</up_axis></asset>", "unterminated"),
    ("<COLLADA><asset></wrong></COLLADA>", "mismatched end tag"),
    ("<scene/>", "not <COLLADA>"),
    ("<COLLADA><asset><up_axis>Y_UP</up_axis></asset></COLLADA>", "instance_visual_scene"),
    ("<COLLADA><asset><up_axis>Y_UP</up_axis></asset><library_visual_scenes><visual_scene id='s'/></library_visual_scenes>"
     "<scene><instance_visual_scene url='#s'/></scene></COLLADA>", "no geometry"),
    ("<COLLADA><asset><up_axis>Y_UP</up_axis></asset><scene><instance_visual_scene url='#nope'/></scene></COLLADA>",
     "unresolved reference"),
])
def test_load_dae_errors(tmp_path, text, needle):
    """Malformed input is an error code + message (the reference exit()s, collada.cpp:117-214)."""
    p = tmp_path / "bad.dae"
    p.write_text(text)
    with pytest.raises(b2rt.B2rtError) as e:
        b2rt.load_dae(str(p))
    assert e.value.code == -5 and needle in str(e.value)


def test_load_dae_missing_file():
    with pytest.raises(b2rt.B2rtError) as e:
        b2rt.load_dae("/nonexistent/scene.dae")
    assert e.value.code == -5 and "cannot open" in str(e.value)


REF_MEDIA = "/root/reference/media/pathtracer"
REF_SCENES = {"CBbunny": "advanced/CBbunny.dae", "CBcoil": "advanced/CBcoil.dae", "CBempty": "advanced/CBempty.dae",
              "CBgems": "advanced/CBgems.dae", "CBspheres": "advanced/CBspheres.dae",
              "CBspheres_lambertian": "advanced/CBspheres_lambertian.dae", "floating": "basic/floating.dae",
              "plane1024": "basic/plane1024.dae", "sphere_diffuse": "basic/sphere_diffuse.dae", "trigs1": "basic/trigs1.dae",
              "trigs5": "basic/trigs5.dae", "trigs10": "basic/trigs10.dae"}


@pytest.mark.skipif(not os.path.isdir(REF_MEDIA), reason="reference media only exists in the build container")
@pytest.mark.parametrize("name", sorted(REF_SCENES))
def test_load_dae_reference_media(name):
    """Every bundled scene under scenes/ is what the C++ loader makes of the reference's own .dae."""
    _same_scene(b2rt.load_dae(os.path.join(REF_MEDIA, REF_SCENES[name])), Scene.load(scene_path(name)))


def test_parallel_build_matches_serial_build():
    """The host builder's worker pool (parallel bounds / binning / partition of large nodes, subtrees as pool items)
    must produce the tree the single-threaded build produces: same subtree / node / leaf / exit counts and blob size."""
    code = ("import sys, json; sys.path.insert(0, %r); import b2rt; from b2rt.scene import Scene, random_soup; "
            "sc = Scene.load(%r); out = [b2rt.validate_bvh_host(sc, 4, 4, 0), b2rt.validate_bvh_host(sc, 2, 8, 0), "
            "b2rt.validate_bvh_host(random_soup(150000), 4, 4, 0)]; print(json.dumps(out))"
            % (os.path.join(ROOT, "cuda-raytracer_b200"), scene_path("CBbunny")))
    res = []
    for threads in ("1", "3", "8"):
        env = dict(os.environ, B2RT_BUILD_THREADS=threads)
        res.append(subprocess.check_output([sys.executable, "-c", code], env=env, text=True).strip().splitlines()[-1])
    assert res[0] == res[1] == res[2], res


def test_bench_workloads_and_strong_scaling_split():
    """bench.py --workload: cfg2 keeps its spp per GPU (weak scaling, the driver's run); cfg3 / cfg4 divide the job's spp
    over the ranks (BASELINE configs[2] / [3]: "spp sharded across 1/2/4/8 B200") and refuse a split that does not divide."""
    sys.path.insert(0, ROOT)
    import bench
    sc, cam, wl = bench.load_workload("cfg2", 0, 8)
    assert (wl["spp"], wl["scaling"], wl["width"], wl["height"], wl["depth"]) == (64, "weak", 1024, 768, 8)
    for world, spp in ((1, 256), (2, 128), (4, 64), (8, 32)):
        sc3, cam3, wl3 = bench.load_workload("cfg3", 0, world)
        assert wl3["spp"] == spp and wl3["spp_job"] == 256 and wl3["scaling"] == "strong" and sc3.n_tris == 114316
    with pytest.raises(SystemExit):
        bench.load_workload("cfg3", 0, 3)
    assert set(bench.BASELINE_INDEX) == set(bench.WORKLOADS)


# ---- host helpers added in round 2 ------------------------------------------------------------------------------
def test_save_png_round_trip(tmp_path):
    from PIL import Image
    rng = np.random.default_rng(5)
    h, w = 37, 53
    px = rng.integers(0, 2 ** 32, size=(h, w), dtype=np.uint64).astype(np.uint32)
    path = str(tmp_path / "t.png")
    b2rt.save_png(path, px)
    img = np.asarray(Image.open(path).convert("RGBA"))
    want = np.stack([(px[::-1] >> s) & 255 for s in (0, 8, 16, 24)], -1).astype(np.uint8)   # top row first, R G B A bytes
    assert np.array_equal(img, want)


def _read_exr_scanlines(path):
    """Minimal reader of the layout b2rt_save_exr writes (OpenEXR 2 scanline, NO_COMPRESSION, FLOAT B G R)."""
    import struct
    d = open(path, "rb").read()
    assert struct.unpack_from("<II", d, 0) == (20000630, 2)
    at = 8
    attrs = {}
    while d[at] != 0:
        e = d.index(b"\0", at); name = d[at:e].decode(); at = e + 1
        e = d.index(b"\0", at); typ = d[at:e].decode(); at = e + 1
        (size,) = struct.unpack_from("<I", d, at); at += 4
        attrs[name] = (typ, d[at:at + size]); at += size
    at += 1
    assert attrs["compression"] == ("compression", b"\0") and attrs["lineOrder"] == ("lineOrder", b"\0")
    chans, c = [], attrs["channels"][1]
    k = 0
    while c[k] != 0:
        e = c.index(b"\0", k); chans.append((c[k:e].decode(), struct.unpack_from("<I", c, e + 1)[0])); k = e + 1 + 16
    assert chans == [("B", 2), ("G", 2), ("R", 2)]
    x0, y0, x1, y1 = struct.unpack("<4i", attrs["dataWindow"][1])
    w, h = x1 - x0 + 1, y1 - y0 + 1
    offs = struct.unpack_from(f"<{h}Q", d, at)
    out = np.zeros((h, w, 3), np.float32)
    for y in range(h):
        yy, nb = struct.unpack_from("<iI", d, offs[y])
        assert yy == y and nb == w * 12
        line = np.frombuffer(d, np.float32, 3 * w, offs[y] + 8).reshape(3, w)
        out[y, :, 2], out[y, :, 1], out[y, :, 0] = line[0], line[1], line[2]
    assert offs[-1] + 8 + w * 12 == len(d)
    return out


def test_save_exr_layout_and_values(tmp_path):
    rng = np.random.default_rng(6)
    rgb = rng.standard_normal((19, 31, 3)).astype(np.float32) * 100
    path = str(tmp_path / "t.exr")
    b2rt.save_exr(path, rgb)
    got = _read_exr_scanlines(path)
    assert np.array_equal(got, rgb[::-1])            # the file stores the top row first
    os.environ["OPENCV_IO_ENABLE_OPENEXR"] = "1"
    try:
        import cv2
        img = cv2.imread(path, cv2.IMREAD_UNCHANGED)
    except Exception:
        img = None
    if img is not None:                              # an independent EXR reader, when this OpenCV build has one
        assert np.array_equal(img[..., ::-1], rgb[::-1])


def test_camera_look_at_reproduces_the_reference_basis():
    """src/cudaRenderer.cu:1592-1599: c_dir = -lookAt, left = (0,1,0) x c_dir, up = left x c_dir; the ray through screen
    coordinates (u, v) is k.x left + k.y up + k.z lookAt with k = (u - .5, -(v - .5), 1) (:347)."""
    o = np.array([0.3, 0.75, 3.0], np.float32)
    L = np.array([0.2, -0.1, -1.0], np.float64); L /= np.linalg.norm(L)
    cam = b2rt.camera_look_at(o, L.astype(np.float32))
    c2w = np.array(cam.c2w[:], np.float64).reshape(3, 3)
    left = np.cross([0, 1, 0], -L); left /= np.linalg.norm(left)
    up = np.cross(left, -L); up /= np.linalg.norm(up)
    np.testing.assert_allclose(c2w[0], left, atol=1e-6)
    np.testing.assert_allclose(c2w[1], -up, atol=1e-6)
    np.testing.assert_allclose(c2w[2], -L, atol=1e-6)
    assert list(cam.pos[:]) == [float(v) for v in o]
    th = np.tan(np.radians(cam.hfov_deg) / 2)
    assert abs(th - 0.5) < 1e-6 and cam.vfov_deg == cam.hfov_deg
    for u, v in ((0.1, 0.8), (0.5, 0.5), (0.9, 0.2)):
        k = np.array([u - .5, -(v - .5), 1.0])
        ref = k[0] * left + k[1] * up + k[2] * L
        ours = c2w[0] * (2 * u - 1) * th + c2w[1] * (2 * v - 1) * th - c2w[2]   # Camera::generate_ray, src/camera.h:71-81
        np.testing.assert_allclose(ours / np.linalg.norm(ours), ref / np.linalg.norm(ref), atol=1e-6)
    with pytest.raises(b2rt.B2rtError):
        b2rt.camera_look_at(o, np.array([0, 1, 0], np.float32))


def test_scene_load_rejects_corrupt_headers(tmp_path):
    good = open(scene_path("CBempty"), "rb").read()
    bad = bytearray(good); bad[8:12] = (0x7FFFFFFF).to_bytes(4, "little")       # n_tris far beyond the file size
    p = tmp_path / "bad.b2s"; p.write_bytes(bytes(bad))
    with pytest.raises(b2rt.B2rtError) as e:
        b2rt.load_scene(str(p))
    assert e.value.code == -5
    bad = bytearray(good); bad[4:8] = (7).to_bytes(4, "little")                 # unknown version
    p.write_bytes(bytes(bad))
    with pytest.raises(b2rt.B2rtError):
        b2rt.load_scene(str(p))
    p.write_bytes(good[: len(good) // 2])                                       # truncated
    with pytest.raises(b2rt.B2rtError):
        b2rt.load_scene(str(p))


def test_comm_api_fails_cleanly_without_a_device():
    if b2rt.device_count() > 0:
        pytest.skip("a CUDA device is visible")
    assert b2rt.Comm.version() >= 0
    with pytest.raises(b2rt.B2rtError):
        b2rt.Comm(2, 0, b"\0" * 128, device=0)


def test_load_dae_accepts_glossy(tmp_path):
    """<glossy> in the CMU462 profile (commented out in the reference's parser, collada.cpp:898-907) loads as
    B2RT_MAT_GLOSSY through the C++ loader and through the independent Python converter alike."""
    import subprocess
    src = open(os.path.join(ROOT, "tests", "golden", "mini_scene.dae")).read()
    assert "<mirror>" in src
    i, j = src.index("<mirror>"), src.index("</mirror>") + len("</mirror>")
    dae = tmp_path / "glossy.dae"
    dae.write_text(src[:i] + "<glossy><reflectance>0.7 0.6 0.5</reflectance><roughness>0.25</roughness></glossy>" + src[j:])
    a = b2rt.load_dae(str(dae))
    g = [m for m in a.materials if m["kind"] == 5]
    assert len(g) == 1 and g[0]["roughness"] == 0.25 and [round(v, 6) for v in g[0]["albedo"]] == [0.7, 0.6, 0.5]
    out = tmp_path / "glossy.b2s"
    subprocess.check_call([sys.executable, os.path.join(ROOT, "tools", "dae2scene.py"), str(dae), str(out)])
    _same_scene(a, Scene.load(str(out)))


def test_ctypes_mirror_matches_the_header_layout(tmp_path):
    """b2rt/_abi.py restates the structs of include/b2rt.h by hand: every field's name, offset and size is compared with
    what the C compiler makes of the header (a field added on one side only shifts everything behind it silently)."""
    import ctypes as C
    import re
    import subprocess
    from b2rt import _abi
    pairs = [("b2rt_material", _abi.Material), ("b2rt_light", _abi.Light), ("b2rt_scene_desc", _abi.SceneDesc),
             ("b2rt_camera", _abi.Camera), ("b2rt_config", _abi.Config), ("b2rt_stats", _abi.Stats),
             ("b2rt_scene_file", _abi.SceneFile)]
    header = open(os.path.join(ROOT, "include", "b2rt.h")).read()
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "b2rt.h"', 'int main(void) {']
    for cname, cls in pairs:
        body = re.search(r"typedef struct %s \{(.*?)\} %s;" % (cname, cname), header, re.S).group(1)
        body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
        c_fields = []
        for decl in body.split(";"):
            if decl.strip():   # "float a[3], b" -> a, b
                parts = decl.split(",")
                names = [parts[0].strip().split()[-1]] + [q.strip() for q in parts[1:]]
                c_fields += [re.sub(r"\[.*\]", "", q).lstrip("*").strip() for q in names]
        assert c_fields == [n for n, _ in cls._fields_], (cname, c_fields)
        lines.append(f'  printf("{cname} %zu\\n", sizeof({cname}));')
        for f in c_fields:
            lines.append(f'  printf("{cname}.{f} %zu %zu\\n", offsetof({cname}, {f}), sizeof((({cname}*)0)->{f}));')
    lines += ['  return 0;', '}']
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), "-o", str(exe), str(src)])
    got = dict((ln.split()[0], [int(x) for x in ln.split()[1:]]) for ln in subprocess.check_output([str(exe)], text=True).splitlines())
    for cname, cls in pairs:
        assert got[cname] == [C.sizeof(cls)], (cname, got[cname], C.sizeof(cls))
        for n, _ in cls._fields_:
            fd = getattr(cls, n)
            assert got[f"{cname}.{n}"] == [fd.offset, fd.size], (cname, n, got[f"{cname}.{n}"], fd.offset, fd.size)


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the CPU arm the driver runs beside the GPU arm): one JSON line on stdout with the
    GPU arm's metric / unit / config keys, `impl: reference`, a `cpu_baseline` describing the run and a zero-copy `e2e`."""
    import json
    import subprocess
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "cfg2", "--cpu-spp", "1",
                          "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, out.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "Mrays/s (all bounces)" and d["unit"] == "Mrays/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 1 and d["value"] > 0
    assert "CBbunny" in d["config"]["workload"] and "BASELINE configs[1]" in d["config"]["workload"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] == (os.cpu_count() or 1) and "1 of 64 spp" in cb["sample"] and cb["value"] == d["value"]
    assert d["e2e"]["value"] == d["value"] and d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0


def test_bench_reference_arm_under_torchrun_prints_once():
    """The driver launches the reference arm like the GPU arm (torchrun for N > 1): rank 0 alone runs and prints the line,
    the other ranks exit 0 without work."""
    import json
    import subprocess
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                          "--master-port", "29541", os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--workload", "cfg2",
                          "--cpu-spp", "1", "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip().startswith("{")]
    assert len(lines) == 1, out.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["n_gpus"] == 2 and d["value"] > 0 and d["scaling"] == "weak"
