"""CPU tests: the oracle against known answers, against the reference's own builder, and against itself
(BVH traversal vs exhaustive search)."""
import json
import os
import subprocess

import numpy as np
import pytest

import orc
from b2rt._abi import Config
from b2rt.scene import Scene, camera_rays, place_camera, random_soup
from conftest import ROOT, scene_path


def test_philox_known_answers():
    # Random123 kat_vectors, philox4x32 10 rounds
    kat = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
           ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
            (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for c, k, out in kat:
        assert tuple(int(x) for x in orc.philox(c, k)) == out


def test_sincos_accuracy():
    for u in np.linspace(0, 0.999999, 2001):
        s, c = orc.sincos2pi(float(u))
        assert abs(s - np.sin(2 * np.pi * u)) < 2e-6 and abs(c - np.cos(2 * np.pi * u)) < 2e-6


def _rays(sc, n, seed):
    rng = np.random.default_rng(seed)
    lo, hi = sc.bbox[:3], sc.bbox[3:]
    o = (lo + (hi - lo) * rng.random((n, 3))).astype(np.float32)
    d = rng.normal(size=(n, 3)); d /= np.linalg.norm(d, axis=1, keepdims=True)
    return o, d.astype(np.float32)


@pytest.mark.parametrize("name,n", [("CBspheres_lambertian", 20000), ("CBgems", 20000), ("CBcoil", 4000), ("CBbunny", 1500),
                                    ("trigs10", 5000), ("sphere_diffuse", 5000)])
def test_bvh_equals_exhaustive(name, n):
    sc = Scene.load(scene_path(name))
    o = orc.OracleScene(sc, 4)
    cam = place_camera(sc, 40, 30)
    ro, rd = camera_rays(cam, 40, 30)
    r2o, r2d = _rays(sc, n, 11)
    org = np.concatenate([ro, r2o]); dirs = np.concatenate([rd, r2d])
    tb, pb = o.intersect(org, dirs, mode="brute")
    tv, pv = o.intersect(org, dirs, mode="bvh")
    assert np.array_equal(pb, pv)
    assert np.array_equal(tb, tv)
    assert (pb != 0xFFFFFFFF).sum() > 0


def test_tie_rule_lowest_prim_id():
    # two coincident triangles: the lower scene index wins in both oracle modes
    tri = np.array([[0, 0, 0, 1, 0, 0, 0, 1, 0]], np.float32)
    sc = Scene(np.concatenate([tri, tri, tri]))
    o = orc.OracleScene(sc, 4)
    org = np.array([[0.2, 0.2, 1.0]], np.float32); d = np.array([[0, 0, -1.0]], np.float32)
    for mode in ("brute", "bvh"):
        t, p = o.intersect(org, d, mode=mode)
        assert p[0] == 0 and t[0] == 1.0


def _ref_dump(scene_file, max_leaf, tmp_path):
    exe = os.path.join(ROOT, "oracle", "_ref", "ref_bvh_dump")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref not built (reference checkout absent)")
    out = tmp_path / "dump.txt"
    subprocess.check_call([exe, scene_file, str(max_leaf), str(out)])
    P, N, L = [], [], []
    for line in open(out):
        t = line.split()
        if t[0] == "P": P.append(int(t[1]))
        elif t[0] == "N": N.append((int(t[1]), int(t[2])))
        elif t[0] == "L": L.append(int(t[1]))
    return P, N, L


@pytest.mark.parametrize("name,max_leaf", [("CBbunny", 32), ("CBbunny", 4), ("CBcoil", 32), ("CBspheres_lambertian", 4),
                                           ("plane1024", 32), ("CBempty", 32)])
def test_builder_matches_reference_builder(name, max_leaf, tmp_path):
    """The restated SAH builder + 4-wide collapse vs the reference's own src/bvh.cpp (oracle/_ref)."""
    P, N, L = _ref_dump(scene_path(name), max_leaf, tmp_path)
    o = orc.OracleScene(Scene.load(scene_path(name)), max_leaf)
    d = o.bvh_dump()
    assert d["order"].tolist() == P
    assert list(zip(d["start"].tolist(), d["range"].tolist())) == N
    assert o.wide_levels() == L


def test_builder_golden_fixture():
    """Same pin, from the committed fixture (tests/golden/ref_bvh.json, written by tools/make_golden.py with the
    reference's builder) so it also holds where oracle/_ref is absent."""
    g = json.load(open(os.path.join(ROOT, "tests", "golden", "ref_bvh.json")))
    for key, ref in g.items():
        name, ml = key.rsplit(":", 1)
        o = orc.OracleScene(Scene.load(scene_path(name)), int(ml))
        d = o.bvh_dump()
        assert len(d["start"]) == ref["binary_nodes"]
        assert o.wide_levels() == ref["wide_levels"]
        assert int(np.sum(d["order"].astype(np.uint64) * (np.arange(len(d["order"]), dtype=np.uint64) % 65521)) % (1 << 61)) == ref["order_checksum"]
        # note: the reference keeps a node as a leaf when no SAH split beats the leaf cost (bvh.cpp:200-203),
        # so leaves may exceed max_leaf_size; every primitive is still in exactly one leaf
        leaves = d["left"] < 0
        assert int(d["range"][leaves].sum()) == len(d["order"])


def test_reference_survey_facts():
    # CBbunny -> 1,891 four-wide nodes, levels 1/4/16/55/207/660/731/217 with max_leaf 32
    o = orc.OracleScene(Scene.load(scene_path("CBbunny")), 32)
    assert o.wide_levels() == [1, 4, 16, 55, 207, 660, 731, 217]
    o = orc.OracleScene(Scene.load(scene_path("CBcoil")), 32)
    assert o.wide_levels() == [1, 4, 16, 59, 215, 217, 34]


def test_render_golden_and_energy():
    sc = Scene.load(scene_path("CBspheres_lambertian"))
    cam = place_camera(sc, 48, 36)
    o = orc.OracleScene(sc, 4)
    img = o.render(cam, Config(ns_aa=4, max_ray_depth=4, ns_area_light=1, seed=1), 48, 36, threads=2)
    g = np.load(os.path.join(ROOT, "tests", "golden", "oracle_cbspheres_48x36_s4_d4.npy"))
    assert np.array_equal(img, g)            # deterministic for any thread count
    img1 = o.render(cam, Config(ns_aa=4, max_ray_depth=4, ns_area_light=1, seed=1), 48, 36, threads=1)
    assert np.array_equal(img, img1)
    assert 0.05 < img.mean() < 0.3
    # sample sharding: two halves of the sample set average to the full set (up to fp32 summation order)
    a = o.render(cam, Config(ns_aa=2, max_ray_depth=4, ns_area_light=1, seed=1, sample_first=0, sample_stride=2), 48, 36)
    b = o.render(cam, Config(ns_aa=2, max_ray_depth=4, ns_area_light=1, seed=1, sample_first=1, sample_stride=2), 48, 36)
    np.testing.assert_allclose(0.5 * (a + b), img, rtol=1e-5, atol=1e-6)


def test_reference_appearance_fixture():
    """Loose appearance pin against the course staff's golden PNG (media/pathtracer/reference_results/sky/CBbunny.png),
    committed as a 64x48 thumbnail by tools/make_golden.py: same camera placement rule, tone map and estimator
    => low-resolution images agree closely."""
    thumb = np.load(os.path.join(ROOT, "tests", "golden", "ref_CBbunny_thumb64x48.npy")).astype(np.float32) / 255.0
    sc = Scene.load(scene_path("CBbunny"))
    w, h = 64, 48
    cam = place_camera(sc, w, h)
    o = orc.OracleScene(sc, 4)
    img = o.render(cam, Config(ns_aa=64, max_ray_depth=4, ns_area_light=2, seed=2), w, h)
    ldr = orc.tonemap(img)[::-1]
    rgb = np.stack([(ldr >> s) & 255 for s in (0, 8, 16)], -1).astype(np.float32) / 255.0
    err = np.sqrt(np.mean((rgb - thumb) ** 2))
    assert err < 0.12, err


def test_soup_bvh_equals_exhaustive():
    sc = random_soup(20000, size=0.05)
    o = orc.OracleScene(sc, 4)
    org, dirs = _rays(sc, 300, 5)
    tb, pb = o.intersect(org, dirs, mode="brute")
    tv, pv = o.intersect(org, dirs, mode="bvh")
    assert np.array_equal(pb, pv) and np.array_equal(tb, tv)


def test_reconstruction_filters_properties():
    """Oracle restatement of the Gaussian / joint bilateral filters (b2rt_config.filter_kind): normalised weights (a
    constant image is a fixed point, also at the border), a very wide range scale turns the bilateral filter into the
    5x5 binomial blur, a narrow one keeps a step edge that the Gaussian smears."""
    rng = np.random.default_rng(4)
    const = np.full((9, 13, 3), 0.37, np.float32)
    for kind in (1, 2):
        np.testing.assert_allclose(orc.recon_filter(const, kind), const, rtol=1e-6)
    img = rng.random((20, 31, 3)).astype(np.float32)
    b = np.array([1, 4, 6, 4, 1], np.float64)
    k5 = np.outer(b, b)
    wide = orc.recon_filter(img, 2, 1e6)
    y, x = 10, 15                                   # interior pixel: plain 5x5 binomial average
    want = (img[y - 2:y + 3, x - 2:x + 3].astype(np.float64) * k5[..., None]).sum((0, 1)) / k5.sum()
    np.testing.assert_allclose(wide[y, x], want, rtol=1e-5)
    step = np.zeros((16, 16, 3), np.float32); step[:, 8:] = 1.0
    g = orc.recon_filter(step, 1); bl = orc.recon_filter(step, 2, 0.05)
    assert 0.2 < g[8, 7, 0] < 0.3 and g[8, 8, 0] > 0.7          # Gaussian: (1,2,1)/4 across the edge
    assert bl[8, 7, 0] < 0.01 and bl[8, 8, 0] > 0.99             # bilateral: the edge survives


def test_atan2_turns_polynomial():
    """The arctangent the environment-map look-up is built on (oracle/oracle.cpp atan2_turns, restated in
    csrc/rt_device.cuh): within 3e-7 turns of libm over all four quadrants, exact on the axes, result in [0, 1)."""
    rng = np.random.default_rng(3)
    pts = rng.standard_normal((4000, 2)).astype(np.float32)
    for y, x in pts:
        got = orc.atan2_turns(y, x)
        want = (np.arctan2(float(y), float(x)) / (2 * np.pi)) % 1.0
        d = abs(got - want)
        assert min(d, 1.0 - d) <= 3e-7 and 0.0 <= got < 1.0
    assert orc.atan2_turns(0.0, 1.0) == 0.0 and orc.atan2_turns(1.0, 0.0) == 0.25
    assert orc.atan2_turns(0.0, -1.0) == 0.5 and orc.atan2_turns(-1.0, 0.0) == 0.75 and orc.atan2_turns(0.0, 0.0) == 0.0


def test_oracle_environment_and_glossy_estimators():
    """Sanity of the two estimators the reference leaves as stubs (EnvironmentLight, GlossyBSDF): a furnace -- a diffuse
    sphere of albedo a under a constant environment of radiance 1 -- converges to the closed form a / (1 - a) * ... per
    bounce sum, and a glossy lobe conserves energy (its hemispherical albedo is <= the reflectance)."""
    from b2rt.scene import Scene, MAT_DIFFUSE, MAT_GLOSSY
    from b2rt._abi import Camera
    a = 0.5
    sc = Scene(np.zeros((0, 9), np.float32), spheres=np.array([[0, 0, 0, 1]], np.float32), sphere_material=np.zeros(1, np.uint32),
               materials=[dict(kind=MAT_DIFFUSE, albedo=(a, a, a))], lights=[])
    o = orc.OracleScene(sc, 4)
    o.set_envmap(np.ones((8, 16, 3), np.float32))
    cam = Camera(); cam.pos[:] = [0, 0, 4]; cam.c2w[:] = [1, 0, 0, 0, 1, 0, 0, 0, 1]; cam.hfov_deg = cam.vfov_deg = 60.0
    depth = 6
    img = o.render(cam, Config(ns_aa=64, max_ray_depth=depth, ns_area_light=1, seed=1), 32, 32)
    # convex object, constant environment: every interaction's NEE sample sees the environment iff it points outward;
    # E[direct] per interaction = albedo (cosine-weighted visible fraction is 1), indirect rays all escape -> radiance = a
    # for depth >= 1 at the first hit plus 0 from the escaped bounce rays (environment counted only after delta bounces)
    centre = img[13:19, 13:19].mean()
    assert abs(centre - a) < 0.04, centre
    corner = img[0, 0]
    assert np.allclose(corner, 1.0)                    # camera rays that miss the sphere see the environment itself
    # glossy: hemispherical albedo by quadrature of f * cos over the hemisphere <= reflectance
    g = Scene(np.zeros((0, 9), np.float32), spheres=np.array([[0, 0, 0, 1]], np.float32), sphere_material=np.zeros(1, np.uint32),
              materials=[dict(kind=MAT_GLOSSY, albedo=(0.8, 0.8, 0.8), roughness=0.2)], lights=[])
    og = orc.OracleScene(g, 4)
    og.set_envmap(np.ones((8, 16, 3), np.float32))
    img = og.render(cam, Config(ns_aa=256, max_ray_depth=2, ns_area_light=1, seed=2), 8, 8)
    assert 0.05 < img[3:5, 3:5].mean() <= 0.8 + 0.1
