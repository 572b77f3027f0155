"""GPU parity tests (run on the B200 box: pytest -m gpu).  Everything goes through the C ABI (libb2rt.so via the
ctypes binding) and is compared with the CPU oracle on the same seeded inputs.

Bars: closest-hit primitive ids and distances, any-hit flags: BIT-EXACT.  Radiance: the integrator's fp32
arithmetic is restated identically in the oracle and in the kernels (explicit fma contract, -fmad=false), so the
HDR frame is expected to be bit-identical; the asserted tolerance is per-pixel RMSE <= 1e-6 and max abs diff <= 1e-5
(north_star: "within a stated per-pixel RMSE tolerance at equal spp").  8-bit tone-mapped output: +-1 LSB (powf)."""
import os

import numpy as np
import pytest

import b2rt
import orc
from b2rt._abi import Config
from b2rt.scene import Scene, camera_rays, place_camera, random_soup, subdivide
from conftest import scene_path

pytestmark = pytest.mark.gpu

MISS = 0xFFFFFFFF


def _rays(sc, n, seed):
    rng = np.random.default_rng(seed)
    lo, hi = sc.bbox[:3], sc.bbox[3:]
    o = (lo + (hi - lo) * rng.random((n, 3))).astype(np.float32)
    d = rng.normal(size=(n, 3)); d /= np.linalg.norm(d, axis=1, keepdims=True)
    return o, d.astype(np.float32)


def _mixed_rays(sc, n_random, seed, w=160, h=120):
    cam = place_camera(sc, w, h)
    ro, rd = camera_rays(cam, w, h)
    r2o, r2d = _rays(sc, n_random, seed)
    return np.concatenate([ro, r2o]), np.concatenate([rd, r2d])


@pytest.mark.parametrize("n_random,w,h", [(0, 1024, 768), (123457, 640, 480), (700001, 1024, 768), (3000077, 160, 120)])
def test_closest_hit_bit_exact_mid_sized_batches(n_random, w, h):
    """Ray counts between 'a tile or two per CTA' and 'thousands of tiles per CTA': the end of the level-0 tile stream,
    where the ring buffers of a CTA claim global tiles out of order (regression: a warp that saw the end marker on one
    buffer stopped while the other buffer still held the last real tile -> lost rays; before that, a hang).  Counts
    that are not multiples of 128 / 4 exercise the partial last tile."""
    sc = Scene.load(scene_path("CBbunny"))
    o = orc.OracleScene(sc, 4)
    bvh = b2rt.BVHAccel(sc)
    org, dirs = _mixed_rays(sc, n_random, 11, w, h)
    for rep in range(2):
        t, p = bvh.intersect(org, dirs)
        tr, pr = o.intersect(org, dirs, mode="bvh")
        assert np.array_equal(p, pr), f"{np.sum(p != pr)} primitive ids differ of {len(org)}"
        assert np.array_equal(t, tr)
    tmax = np.full(len(org), 2.5, np.float32)
    occ = bvh.occluded(org, dirs, None, tmax)
    _, pw = o.intersect(org, dirs, None, tmax)
    assert np.array_equal(occ, pw != MISS)       # any hit inside the window <=> a closest hit exists
    bvh.close()


@pytest.mark.parametrize("name", ["CBbunny", "CBcoil", "CBgems", "CBspheres_lambertian", "CBspheres", "CBempty", "trigs10",
                                  "sphere_diffuse", "plane1024", "floating"])
@pytest.mark.parametrize("width,treelet_bytes,max_leaf", [(4, 0, 4), (8, 0, 4), (4, 8192, 2), (8, 100000, 8), (2, 0, 4), (16, 0, 8),
                                                          (2, 8192, 16), (16, 65536, 32)])
def test_closest_hit_bit_exact(name, width, treelet_bytes, max_leaf):
    sc = Scene.load(scene_path(name))
    o = orc.OracleScene(sc, 4)
    bvh = b2rt.BVHAccel(sc, max_leaf_size=max_leaf, width=width, treelet_bytes=treelet_bytes)
    org, dirs = _mixed_rays(sc, 40000, 3)
    t, p = bvh.intersect(org, dirs)
    tr, pr = o.intersect(org, dirs, mode="bvh")
    assert np.array_equal(p, pr), f"{np.sum(p != pr)} primitive ids differ"
    assert np.array_equal(t, tr)
    st = bvh.stats()
    assert st["kernel_launches"] >= 1 and st["subtree_visits"] >= len(org)
    # any hit on every scene / width / subtree budget: occluded inside a random window <=> the oracle's closest hit
    # restricted to that window exists
    rng = np.random.default_rng(17)
    bb = bvh.get_bbox()
    diag = float(np.linalg.norm(bb[3:] - bb[:3]))
    tmin = (rng.random(len(org)) * 0.1 * diag).astype(np.float32)
    tmax = (tmin + rng.random(len(org)) * 0.6 * diag).astype(np.float32)
    occ = bvh.occluded(org, dirs, tmin, tmax)
    _, pw = o.intersect(org, dirs, tmin, tmax)
    assert np.array_equal(occ, pw != MISS), f"{int(np.sum(occ != (pw != MISS)))} any-hit flags differ"
    bvh.close()


def test_closest_hit_vs_exhaustive_search():
    """Against the argmin over ALL primitives (what an exact closest hit is, independent of any BVH)."""
    sc = Scene.load(scene_path("CBcoil"))
    o = orc.OracleScene(sc, 4)
    bvh = b2rt.BVHAccel(sc)
    org, dirs = _mixed_rays(sc, 4000, 5, 64, 48)
    t, p = bvh.intersect(org, dirs)
    tr, pr = o.intersect(org, dirs, mode="brute")
    assert np.array_equal(p, pr) and np.array_equal(t, tr)


def test_tmin_tmax_windows_and_any_hit():
    sc = Scene.load(scene_path("CBbunny"))
    o = orc.OracleScene(sc, 4)
    bvh = b2rt.BVHAccel(sc)
    org, dirs = _mixed_rays(sc, 30000, 8)
    rng = np.random.default_rng(1)
    tmin = (rng.random(len(org)) * 0.5).astype(np.float32)
    tmax = (tmin + rng.random(len(org)) * 3).astype(np.float32)
    t, p = bvh.intersect(org, dirs, tmin, tmax)
    tr, pr = o.intersect(org, dirs, tmin, tmax)
    assert np.array_equal(p, pr) and np.array_equal(t, tr)
    hit = p != MISS
    assert np.all((t[hit] >= tmin[hit]) & (t[hit] <= tmax[hit]))
    occ = bvh.occluded(org, dirs, tmin, tmax)
    assert np.array_equal(occ, pr != MISS)       # any hit inside the window <=> a closest hit exists


def test_edge_cases():
    sc = Scene.load(scene_path("CBspheres_lambertian"))
    o = orc.OracleScene(sc, 4)
    bvh = b2rt.BVHAccel(sc)
    # empty batch
    t, p = bvh.intersect(np.zeros((0, 3), np.float32), np.zeros((0, 3), np.float32))
    assert len(t) == 0 and len(p) == 0
    # axis-aligned directions (zero components), rays starting on surfaces, rays leaving the scene
    org = np.array([[0, 0.75, 3], [0, 0.75, 0], [0, 0, 0], [0.5, 1.0, 0.2], [0, 5, 0], [0, 0.75, 0]], np.float32)
    dirs = np.array([[0, 0, -1], [0, -1, 0], [0, 1, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1]], np.float32)
    t, p = bvh.intersect(org, dirs)
    tr, pr = o.intersect(org, dirs)
    assert np.array_equal(p, pr) and np.array_equal(t, tr)
    assert p[4] == MISS and np.isinf(t[4])
    # ragged batch sizes around warp / chunk boundaries
    for n in (1, 31, 33, 1023, 1025, 4097):
        ro, rd = _rays(sc, n, n)
        t, p = bvh.intersect(ro, rd)
        tr, pr = o.intersect(ro, rd)
        assert np.array_equal(p, pr) and np.array_equal(t, tr)


def test_tie_rule_lowest_prim_id():
    tri = np.array([[0, 0, 0, 1, 0, 0, 0, 1, 0]], np.float32)
    quad2 = np.array([[1, 0, 0, 1, 1, 0, 0, 1, 0]], np.float32)     # shares the diagonal edge
    sc = Scene(np.concatenate([tri, tri, quad2, tri]))
    o = orc.OracleScene(sc, 4)
    bvh = b2rt.BVHAccel(sc, max_leaf_size=1)
    n = 2000
    s = np.linspace(0.001, 0.999, n, dtype=np.float32)
    org = np.stack([s, 1 - s, np.ones(n, np.float32)], 1)   # along the shared diagonal
    dirs = np.tile(np.array([[0, 0, -1]], np.float32), (n, 1))
    t, p = bvh.intersect(org, dirs)
    tr, pr = o.intersect(org, dirs, mode="brute")
    assert np.array_equal(p, pr) and np.array_equal(t, tr)
    assert set(np.unique(p)) <= {0, 2}


def test_soup_parity_and_multilevel():
    sc = random_soup(300000, size=0.02)
    o = orc.OracleScene(sc, 4)
    bvh = b2rt.BVHAccel(sc, treelet_bytes=16384)
    assert bvh.stats()["bvh_levels"] >= 3
    org, dirs = _rays(sc, 60000, 21)
    t, p = bvh.intersect(org, dirs)
    tr, pr = o.intersect(org, dirs)
    assert np.array_equal(p, pr) and np.array_equal(t, tr)
    assert (p != MISS).mean() > 0.5


@pytest.mark.parametrize("first,growth,passes", [(0.05, 2.0, 3), (0.3, 4.0, 4), (1e-3, 8.0, 6), (10.0, 4.0, 2)])
def test_distance_sliced_trace_is_identical(first, growth, passes):
    """b2rt_bvh_set_slicing: tracing the rays slice by slice (front-to-back order across subtrees) must return the same
    (t, prim) argmin and the same any-hit flags as the oracle, whatever the slice lengths -- including windows
    [tmin, tmax], rays that start outside the scene box, rays that miss it, and slices far longer than the scene."""
    sc = Scene.load(scene_path("CBbunny"))
    o = orc.OracleScene(sc, 4)
    bvh = b2rt.BVHAccel(sc, treelet_bytes=4096, max_leaf_size=2)
    assert bvh.stats()["bvh_levels"] >= 3
    bvh.set_slicing(first, growth, passes)
    org, dirs = _mixed_rays(sc, 150001, 5)
    far = _rays(sc, 5000, 6)
    org = np.concatenate([org, far[0] * 3 + 5]); dirs = np.concatenate([dirs, -far[1]])     # from outside, mostly missing
    t, p = bvh.intersect(org, dirs)
    tr, pr = o.intersect(org, dirs, mode="bvh")
    assert np.array_equal(p, pr), f"{np.sum(p != pr)} primitive ids differ of {len(org)}"
    assert np.array_equal(t, tr)
    rng = np.random.default_rng(2)
    tmin = (rng.random(len(org)) * 0.5).astype(np.float32)
    tmax = (tmin + rng.random(len(org)) * 3).astype(np.float32)
    t, p = bvh.intersect(org, dirs, tmin, tmax)
    tr, pr = o.intersect(org, dirs, tmin, tmax)
    assert np.array_equal(p, pr) and np.array_equal(t, tr)
    occ = bvh.occluded(org, dirs, tmin, tmax)
    assert np.array_equal(occ, pr != MISS)
    bvh.close()


def test_degenerate_rays_sliced_and_plain():
    """Empty windows (tmin > tmax), zero and NaN directions, origins far outside the scene, infinite tmax: the sliced
    and the plain trace agree with the oracle (a NaN never hits: every accept test is written positively)."""
    sc = Scene.load(scene_path("CBbunny"))
    o = orc.OracleScene(sc, 4)
    bvh = b2rt.BVHAccel(sc, treelet_bytes=4096, max_leaf_size=2)
    org, dirs = _mixed_rays(sc, 4000, 9, 64, 48)
    n = len(org)
    tmin = np.zeros(n, np.float32); tmax = np.full(n, np.inf, np.float32)
    tmin[0::7] = 2.0; tmax[0::7] = 1.0                      # empty windows
    dirs[1::11] = 0.0                                        # zero directions
    dirs[2::13, 1] = np.nan                                  # NaN component
    org[3::17] = org[3::17] * 1e4 + 3e4                      # far away
    org[4::19, 0] = np.inf                                   # infinite origin component
    tmax[5::23] = 0.0                                        # zero-length window at the origin
    tr, pr = o.intersect(org, dirs, tmin, tmax)
    for first in (0.0, 0.2, 5.0):
        bvh.set_slicing(first, 3.0, 3)
        t, p = bvh.intersect(org, dirs, tmin, tmax)
        assert np.array_equal(p, pr), (first, int(np.sum(p != pr)))
        assert np.array_equal(t, tr)
        occ = bvh.occluded(org, dirs, tmin, tmax)
        assert np.array_equal(occ, pr != MISS)
    bvh.close()


def test_soup_automatic_slicing_parity_and_fewer_visits():
    """On a deep subtree graph the automatic slices (2 mean free paths) are on by default: same hits as the oracle and
    as the unsliced trace, with fewer subtree visits per ray."""
    sc = random_soup(300000, size=0.05)     # mean free path ~0.02: the automatic first slice (2 mfp) is far below diagonal / 8
    o = orc.OracleScene(sc, 4)
    bvh = b2rt.BVHAccel(sc, treelet_bytes=16384)
    assert bvh.stats()["bvh_levels"] >= 3
    org, dirs = _rays(sc, 200000, 22)
    t, p = bvh.intersect(org, dirs)
    v_sliced = bvh.stats()["subtree_visits"]
    bvh.set_slicing(0.0)
    t0, p0 = bvh.intersect(org, dirs)
    v_plain = bvh.stats()["subtree_visits"]
    assert np.array_equal(p, p0) and np.array_equal(t, t0)
    tr, pr = o.intersect(org[:50000], dirs[:50000])
    assert np.array_equal(p[:50000], pr) and np.array_equal(t[:50000], tr)
    assert v_sliced < 0.8 * v_plain, (v_sliced, v_plain)
    bvh.close()


@pytest.mark.parametrize("name", ["CBbunny", "CBcoil", "CBspheres", "trigs1", "trigs10", "sphere_diffuse", "plane1024"])
@pytest.mark.parametrize("width,treelet_bytes,max_leaf", [(4, 0, 4), (8, 0, 4), (4, 4096, 2), (4, 65536, 8)])
def test_device_built_bvh_structure_and_parity(name, width, treelet_bytes, max_leaf):
    """b2rt_bvh_build_device (LBVH on the GPU): the blob passes the structural validator (every primitive stored once
    with its exact record, child boxes contain their contents, exits point one level down, byte / node / stack budgets)
    and closest hits are bit-identical to the oracle -- an exact closest hit does not depend on which BVH culled."""
    sc = Scene.load(scene_path(name))
    bvh = b2rt.BVHAccel(sc, max_leaf_size=max_leaf, width=width, treelet_bytes=treelet_bytes, builder="gpu")
    v = bvh.validate()
    st = bvh.stats()
    assert v["subtrees"] == st["bvh_subtrees"] and v["levels"] == st["bvh_levels"] and v["wide_nodes"] == st["bvh_nodes"]
    assert np.allclose(bvh.get_bbox(), sc.bbox)
    o = orc.OracleScene(sc, 4)
    org, dirs = _mixed_rays(sc, 20000, 3, 96, 72)
    t, p = bvh.intersect(org, dirs)
    tr, pr = o.intersect(org, dirs, mode="bvh")
    assert np.array_equal(p, pr), f"{np.sum(p != pr)} primitive ids differ"
    assert np.array_equal(t, tr)
    bvh.close()


def test_device_built_bvh_soup():
    """300 K triangle soup: device build validated structurally, same hits as the host-built BVH and as the oracle."""
    sc = random_soup(300000, size=0.02)
    g = b2rt.BVHAccel(sc, treelet_bytes=16384, builder="gpu")
    v = g.validate()
    assert v["levels"] >= 3 and v["leaves"] >= 300000 // 4
    h = b2rt.BVHAccel(sc, treelet_bytes=16384)
    org, dirs = _rays(sc, 200000, 23)
    tg, pg = g.intersect(org, dirs)
    th, ph = h.intersect(org, dirs)
    assert np.array_equal(pg, ph) and np.array_equal(tg, th)
    o = orc.OracleScene(sc, 4)
    tr, pr = o.intersect(org[:30000], dirs[:30000])
    assert np.array_equal(pg[:30000], pr) and np.array_equal(tg[:30000], tr)
    g.close(); h.close()


RENDER_CASES = [
    # scene, w, h, spp, depth, ns_area_light
    ("CBspheres_lambertian", 480, 360, 16, 4, 1),   # BASELINE configs[0] at full size
    ("CBbunny", 256, 192, 4, 8, 1),                 # configs[1] knobs at reduced size (full size below, via properties)
    ("CBgems", 160, 120, 8, 6, 2),                  # glass
    ("CBcoil", 160, 120, 4, 5, 1),                  # mirror
    ("CBspheres", 160, 120, 8, 6, 1),               # glass + mirror spheres
    ("CBempty", 96, 72, 3, 2, 4),
    ("floating", 96, 72, 2, 3, 1),
]


@pytest.mark.parametrize("name,w,h,spp,depth,nsl", RENDER_CASES)
def test_radiance_parity(name, w, h, spp, depth, nsl):
    sc = Scene.load(scene_path(name))
    cam = place_camera(sc, w, h)
    pt = b2rt.PathTracer(ns_aa=spp, max_ray_depth=depth, ns_area_light=nsl, seed=11)
    pt.set_scene(sc); pt.set_camera(cam); pt.set_frame_size(w, h)
    assert pt.start_raytracing()
    pt.wait()
    assert pt.is_done()
    img = pt.hdr()
    o = orc.OracleScene(sc, 4)
    ref = o.render(cam, Config(ns_aa=spp, max_ray_depth=depth, ns_area_light=nsl, seed=11), w, h)
    rmse = float(np.sqrt(np.mean((img - ref) ** 2)))
    assert rmse <= 1e-6, rmse
    assert float(np.abs(img - ref).max()) <= 1e-5
    st = pt.stats()
    cs = o.last_stats
    assert (st["rays_camera"], st["rays_bounce"], st["rays_shadow"]) == (cs["rays_camera"], cs["rays_bounce"], cs["rays_shadow"])
    # tone-mapped 8-bit output: toColor formula, +-1 LSB
    ldr = pt.ldr()
    ref_ldr = orc.tonemap(ref)
    for s in (0, 8, 16, 24):
        d = np.abs(((ldr >> s) & 255).astype(np.int32) - ((ref_ldr >> s) & 255).astype(np.int32))
        assert d.max() <= 1
    pt.close()


def test_waves_and_width_do_not_change_the_image():
    sc = Scene.load(scene_path("CBbunny"))
    w, h = 128, 96
    cam = place_camera(sc, w, h)
    imgs = []
    for kw in (dict(), dict(max_wave_paths=5000), dict(bvh_width=8, treelet_bytes=20000), dict(max_leaf_size=8, max_wave_paths=40000)):
        pt = b2rt.PathTracer(ns_aa=6, max_ray_depth=4, ns_area_light=1, seed=2, **kw)
        pt.set_scene(sc); pt.set_camera(cam); pt.set_frame_size(w, h)
        pt.render()
        imgs.append(pt.hdr())
        pt.close()
    for im in imgs[1:]:
        assert np.array_equal(im, imgs[0])


def test_full_size_properties_cfg2():
    """BASELINE configs[1] at full size (CBbunny 1024x768, 64 spp, depth 8): too slow for the oracle, so checked through
    size-independent properties: determinism, sample-shard linearity (the multi-GPU decomposition), agreement of a
    sub-window of samples with the oracle, ray-count conservation."""
    sc = Scene.load(scene_path("CBbunny"))
    w, h, spp, depth = 1024, 768, 64, 8
    cam = place_camera(sc, w, h)
    pt = b2rt.PathTracer(ns_aa=spp, max_ray_depth=depth, ns_area_light=1, seed=1)
    pt.set_scene(sc); pt.set_camera(cam); pt.set_frame_size(w, h)
    pt.render(); full = pt.hdr(); st = pt.stats()
    assert st["rays_camera"] == w * h * spp
    assert st["rays_bounce"] < st["rays_camera"] * (depth - 1) and st["rays_shadow"] < st["rays_camera"] * depth
    pt.clear(); pt.render()
    assert np.array_equal(pt.hdr(), full)                                  # deterministic
    halves = []
    for r in range(2):
        pt.set_config(ns_aa=spp // 2, sample_first=r, sample_stride=2)
        pt.clear(); pt.render(); halves.append(pt.hdr())
    np.testing.assert_allclose(0.5 * (halves[0] + halves[1]), full, rtol=1e-4, atol=1e-5)   # linearity over sample shards
    # oracle on a 1-sample shard (sample index 37 of the same global stream)
    pt.set_config(ns_aa=1, sample_first=37, sample_stride=64)
    pt.clear(); pt.render(); one = pt.hdr()
    ref = orc.OracleScene(sc, 4).render(cam, Config(ns_aa=1, max_ray_depth=depth, ns_area_light=1, seed=1, sample_first=37,
                                                    sample_stride=64), w, h)
    assert float(np.sqrt(np.mean((one - ref) ** 2))) <= 1e-6
    assert 0.05 < full.mean() < 0.3
    pt.close()


def test_dragon_class_standin_parity():
    """cfg3 stand-in: CBbunny with the bunny mesh subdivided once, 114,316 triangles."""
    base = Scene.load(scene_path("CBbunny"))
    sc = subdivide(base, 1, select=lambda tv, tm: tm == tm[np.argmax(np.bincount(tm))])
    o = orc.OracleScene(sc, 4)
    bvh = b2rt.BVHAccel(sc)
    org, dirs = _mixed_rays(sc, 50000, 4, 320, 240)
    t, p = bvh.intersect(org, dirs)
    tr, pr = o.intersect(org, dirs)
    assert np.array_equal(p, pr) and np.array_equal(t, tr)
    bvh.close()
    w, h = 192, 108
    cam = place_camera(sc, w, h)
    pt = b2rt.PathTracer(ns_aa=2, max_ray_depth=8, ns_area_light=1, seed=4)
    pt.set_scene(sc); pt.set_camera(cam); pt.set_frame_size(w, h); pt.render()
    ref = o.render(cam, Config(ns_aa=2, max_ray_depth=8, ns_area_light=1, seed=4), w, h)
    assert float(np.sqrt(np.mean((pt.hdr() - ref) ** 2))) <= 1e-6


def test_full_size_properties_cfg3():
    """BASELINE configs[2] at full size (the dragon-class stand-in, 1920x1080, 256 spp, depth 8 -- bench.py's default
    workload, BVH built on the device like there): determinism, linearity over two sample shards (the multi-GPU
    decomposition), a one-sample shard of the same global stream against the oracle at full resolution, ray-count
    conservation; the third frame of the same signature is replayed from a CUDA graph and must not change a bit."""
    from b2rt.scene import cfg3_standin
    sc = cfg3_standin(Scene.load(scene_path("CBbunny")))
    w, h, spp, depth = 1920, 1080, 256, 8
    cam = place_camera(sc, w, h)
    pt = b2rt.PathTracer(ns_aa=spp, max_ray_depth=depth, ns_area_light=1, seed=1)
    pt.set_scene(sc); pt.set_camera(cam); pt.set_frame_size(w, h)
    pt.render(); full = pt.hdr(); st = pt.stats()
    assert st["rays_camera"] == w * h * spp
    assert st["rays_bounce"] < st["rays_camera"] * (depth - 1) and st["rays_shadow"] < st["rays_camera"] * depth
    for _ in range(2):
        pt.clear(); pt.render()
        assert np.array_equal(pt.hdr(), full)                              # deterministic, eager and replayed
    assert pt.stats()["graph_replays"] >= 1
    halves = []
    for r in range(2):
        pt.set_config(ns_aa=spp // 2, sample_first=r, sample_stride=2)
        pt.clear(); pt.render(); halves.append(pt.hdr())
    np.testing.assert_allclose(0.5 * (halves[0] + halves[1]), full, rtol=1e-4, atol=1e-5)
    pt.set_config(ns_aa=1, sample_first=201, sample_stride=256)
    pt.clear(); pt.render(); one = pt.hdr()
    ref = orc.OracleScene(sc, 4).render(cam, Config(ns_aa=1, max_ray_depth=depth, ns_area_light=1, seed=1, sample_first=201,
                                                    sample_stride=256), w, h)
    assert float(np.sqrt(np.mean((one - ref) ** 2))) <= 1e-6
    assert float(np.abs(one - ref).max()) <= 1e-5
    pt.close()


@pytest.mark.parametrize("name", ["CBbunny", "CBspheres", "CBcoil"])
def test_renderer_on_device_built_bvh(name):
    """b2rt_config.bvh_builder = 2: b2rt_set_scene builds the BVH on the device; the frame stays identical to the oracle
    (and therefore to the host-built path)."""
    sc = Scene.load(scene_path(name))
    w, h = 128, 96
    cam = place_camera(sc, w, h)
    cfg = dict(ns_aa=4, max_ray_depth=6, ns_area_light=1, seed=5)
    pt = b2rt.PathTracer(bvh_builder=2, **cfg)
    pt.set_scene(sc); pt.set_camera(cam); pt.set_frame_size(w, h)
    pt.render()
    img = pt.hdr()
    ref = orc.OracleScene(sc, 4).render(cam, Config(**cfg), w, h)
    rmse = float(np.sqrt(np.mean((img - ref) ** 2)))
    assert rmse <= 1e-6 and float(np.abs(img - ref).max()) <= 1e-5, rmse
    pt.set_scene(sc)                 # rebuild into the same device buffers
    pt.clear(); pt.render()
    assert np.array_equal(pt.hdr(), img)
    pt.close()


def test_renderer_slices_dense_deep_scenes_automatically():
    """A lit triangle soup through the path tracer: the subtree graph is deep and the geometry dense, so b2rt_set_scene
    turns the distance slices on by the same rule as b2rt_bvh_build (4 passes per trace); the frame stays identical to
    the oracle's."""
    soup = random_soup(120000, size=0.05)
    sc = Scene(soup.tri_verts, materials=[dict(kind=0, albedo=(0.7, 0.6, 0.5))],
               lights=[dict(kind=1, radiance=(3.0, 3.0, 3.0), position=(0.5, 0.6, 2.0))], cam_dir=(0, 0, 1))
    w, h = 64, 48
    cam = place_camera(sc, w, h)
    cfg = dict(ns_aa=2, max_ray_depth=3, ns_area_light=1, seed=4)
    pt = b2rt.PathTracer(treelet_bytes=8192, **cfg)
    pt.set_scene(sc); pt.set_camera(cam); pt.set_frame_size(w, h)
    pt.render()
    img = pt.hdr()
    st = pt.stats()
    assert st["bvh_levels"] >= 3
    assert st["traverse_launches"] == 4 * st["bvh_levels"] * 2 * 3      # 4 slice passes x levels x (closest + shadow) x depth
    ref = orc.OracleScene(sc, 4).render(cam, Config(**cfg), w, h)
    assert img.max() > 0
    rmse = float(np.sqrt(np.mean((img - ref) ** 2)))
    assert rmse <= 1e-6 and float(np.abs(img - ref).max()) <= 1e-5, rmse
    pt.close()


def test_cfg4_standin_material_mix_parity():
    """BASELINE configs[3] stand-in (glass mesh + mirror spheres in the Cornell box) at test size, mesh not subdivided:
    HDR frame identical to the oracle."""
    from b2rt.scene import cfg4_standin
    sc = cfg4_standin(Scene.load(scene_path("CBbunny")), levels=0)
    w, h = 128, 96
    cam = place_camera(sc, w, h)
    cfg = dict(ns_aa=4, max_ray_depth=8, ns_area_light=1, seed=9)
    pt = b2rt.PathTracer(**cfg)
    pt.set_scene(sc); pt.set_camera(cam); pt.set_frame_size(w, h)
    pt.render()
    img = pt.hdr()
    ref = orc.OracleScene(sc, 4).render(cam, Config(**cfg), w, h)
    assert np.isfinite(img).all()
    rmse = float(np.sqrt(np.mean((img - ref) ** 2)))
    assert rmse <= 1e-6 and float(np.abs(img - ref).max()) <= 1e-5, rmse
    pt.close()


def test_cfg4_standin_full_mesh_parity():
    """BASELINE configs[3] stand-in with the mesh at its FULL size (bunny subdivided twice, 457,228 glass triangles, two
    mirror spheres; device-built BVH) at a small frame: HDR frame identical to the oracle's (own binary SAH BVH)."""
    from b2rt.scene import cfg4_standin
    sc = cfg4_standin(Scene.load(scene_path("CBbunny")), levels=2)
    assert sc.n_tris > 450000
    w, h = 96, 72
    cam = place_camera(sc, w, h)
    cfg = dict(ns_aa=2, max_ray_depth=8, ns_area_light=1, seed=13)
    pt = b2rt.PathTracer(**cfg)
    pt.set_scene(sc); pt.set_camera(cam); pt.set_frame_size(w, h)
    pt.render()
    img = pt.hdr()
    ref = orc.OracleScene(sc, 4).render(cam, Config(**cfg), w, h)
    assert np.isfinite(img).all() and img.max() > 0
    rmse = float(np.sqrt(np.mean((img - ref) ** 2)))
    assert rmse <= 1e-6 and float(np.abs(img - ref).max()) <= 1e-5, rmse
    pt.close()


def test_progressive_frames_replayed_from_a_graph_match_oracle(monkeypatch):
    """CudaRenderer.render() adds samples_per_frame samples per call; from the third call on the frame is replayed from a
    captured CUDA graph with only the first sample index changed (a device word).  Five calls = the oracle's 10-spp
    frame, and the same with graphs disabled, bit for bit."""
    sc = Scene.load(scene_path("CBspheres_lambertian"))
    w, h = 96, 72
    cam = place_camera(sc, w, h)
    imgs = []
    for graphs in (True, False):
        if not graphs:
            monkeypatch.setenv("B2RT_GRAPH", "0")
        r = b2rt.CudaRenderer(samples_per_frame=2, max_ray_depth=4, ns_area_light=1, median_threshold=0, seed=21)
        r.allocOutputImage(w, h); r.loadScene(sc); r.setViewpoint(cam)
        for _ in range(5):
            r.render()
        st = r.pt.stats()
        assert (st["graph_replays"] == 3) if graphs else (st["graph_replays"] == 0), st
        imgs.append(r.pt.hdr())
        r.pt.close()
    ref = orc.OracleScene(sc, 4).render(cam, Config(ns_aa=10, max_ray_depth=4, ns_area_light=1, seed=21), w, h)
    assert np.array_equal(imgs[0], imgs[1])
    rmse = float(np.sqrt(np.mean((imgs[0] - ref) ** 2)))
    assert rmse <= 1e-6 and float(np.abs(imgs[0] - ref).max()) <= 1e-5, rmse


def test_random_scenes_and_configurations_match_oracle():
    """tools/fuzz_frames.py: random scenes (all material kinds, vertex normals, spheres, the three light kinds, optional
    environment map) x random renderer configurations (spp, depth, light samples, wave size, BVH width / leaf / subtree
    budget, host or device builder): every HDR frame identical to the oracle's."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "tools", "fuzz_frames.py"), "12", "77"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert "mismatches: 0" in r.stdout


def test_median_filter_and_progressive_renderer():
    sc = Scene.load(scene_path("CBspheres_lambertian"))
    w, h = 100, 75
    r = b2rt.CudaRenderer(samples_per_frame=2, max_ray_depth=3, ns_area_light=2, median_threshold=32, seed=6)
    r.allocOutputImage(w, h); r.loadScene(sc); r.setup()
    r.render()
    img1 = r.getImage()
    o = orc.OracleScene(sc, 4)
    cam = place_camera(sc, w, h)
    ref1 = o.render(cam, Config(ns_aa=2, max_ray_depth=3, ns_area_light=2, seed=6), w, h)
    np.testing.assert_array_equal(img1[..., :3], orc.median3x3(ref1))       # below the threshold: 3x3 median, border 1.0
    assert np.all(img1[..., 3] == 1.0)
    r.render()                                                              # second frame: samples 2,3 accumulate
    ref2 = o.render(cam, Config(ns_aa=4, max_ray_depth=3, ns_area_light=2, seed=6), w, h)
    np.testing.assert_allclose(r.getImage()[..., :3], orc.median3x3(ref2), rtol=1e-5, atol=1e-6)
    r.setViewpoint(cam)                                                     # resets accumulation
    r.render()
    np.testing.assert_array_equal(r.getImage()[..., :3], orc.median3x3(ref1))


@pytest.mark.parametrize("kind,sigma_r", [(1, 0.0), (2, 0.0), (2, 0.05)])
def test_gaussian_and_bilateral_reconstruction_filters(kind, sigma_r):
    """b2rt_config.filter_kind 1 / 2: the filtered frame is bit-identical to the oracle's restatement
    applied to the oracle's frame, on an image size that is not a multiple of the 32 x 8 tile."""
    sc = Scene.load(scene_path("CBspheres_lambertian"))
    w, h = 101, 75
    r = b2rt.CudaRenderer(samples_per_frame=2, max_ray_depth=3, ns_area_light=2, median_threshold=32, seed=6,
                          filter_kind=kind, filter_sigma_r=sigma_r)
    r.allocOutputImage(w, h); r.loadScene(sc); r.setup()
    r.render()
    img = r.getImage()
    cam = place_camera(sc, w, h)
    ref = orc.OracleScene(sc, 4).render(cam, Config(ns_aa=2, max_ray_depth=3, ns_area_light=2, seed=6), w, h)
    np.testing.assert_array_equal(img[..., :3], orc.recon_filter(ref, kind, sigma_r))
    assert np.all(img[..., 3] == 1.0)
    assert not np.array_equal(img[..., :3], ref)


def test_state_machine_and_errors():
    sc = Scene.load(scene_path("CBempty"))
    pt = b2rt.PathTracer(ns_aa=1)
    assert pt.state == pt.INIT and not pt.start_raytracing()               # only from READY (pathtracer.cpp:184)
    with pytest.raises(b2rt.B2rtError):
        b2rt._check(b2rt.lib().b2rt_start(pt._h))                          # C ABI: error code instead of a silent no-op
    pt.set_scene(sc); pt.set_camera(place_camera(sc, 64, 48)); pt.set_frame_size(64, 48)
    assert pt.state == pt.READY
    pt.set_config(ns_aa=64, max_ray_depth=8)
    assert pt.start_raytracing() and pt.state == pt.RENDERING
    pt.stop()
    assert pt.state == pt.READY
    pt.set_config(ns_aa=1)
    pt.set_frame_size(32, 24)                                              # resize invalidates the frame
    assert pt.start_raytracing(); pt.wait(); assert pt.is_done()
    assert pt.hdr().shape == (24, 32, 3)
    pt.increase_area_light_sample_count(); assert pt.cfg.ns_area_light == 2
    pt.decrease_area_light_sample_count(); assert pt.cfg.ns_area_light == 1
    with pytest.raises(b2rt.B2rtError):
        pt.set_frame_size(0, 10)
    pt.close()


def test_two_gpu_style_sharding_on_one_device():
    """The multi-GPU decomposition emulated on one GPU: two handles render disjoint sample shards, their accumulation
    buffers are summed (what the NCCL reduce does) and resolved."""
    import torch
    from b2rt.dist import resolve_mean, shard_samples
    sc = Scene.load(scene_path("CBcoil"))
    w, h, spp = 96, 72, 8
    cam = place_camera(sc, w, h)
    total = None
    for r in range(2):
        first, stride, cnt = shard_samples(spp, r, 2)
        pt = b2rt.PathTracer(ns_aa=cnt, max_ray_depth=4, seed=13, sample_first=first, sample_stride=stride)
        pt.set_scene(sc); pt.set_camera(cam); pt.set_frame_size(w, h); pt.render()
        acc = pt.accum_tensor().clone()
        total = acc if total is None else total + acc
        pt.close()
    got = resolve_mean(total).cpu().numpy().reshape(h, w, 3)
    ref = orc.OracleScene(sc, 4).render(cam, Config(ns_aa=spp, max_ray_depth=4, ns_area_light=1, seed=13), w, h)
    np.testing.assert_allclose(got, ref, rtol=2e-5, atol=1e-6)


def test_cpp_host_example(tmp_path):
    """The C++ host program (shim classes over the C ABI) renders the same RGBA8 frame as the Python binding."""
    import os
    import subprocess
    from PIL import Image
    from conftest import ROOT
    exe = os.path.join(ROOT, "cuda-raytracer_b200", "examples", "render_scene")
    out = tmp_path / "cpp.png"
    r = subprocess.run([exe, "-s", "4", "-m", "3", "-l", "1", "-r", "96x72", "-w", str(out), scene_path("CBspheres_lambertian")],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    sc = Scene.load(scene_path("CBspheres_lambertian"))
    pt = b2rt.PathTracer(ns_aa=4, max_ray_depth=3, ns_area_light=1)
    pt.set_scene(sc); pt.set_camera(place_camera(sc, 96, 72)); pt.set_frame_size(96, 72); pt.render()
    ldr = pt.ldr()[::-1]
    ref = np.stack([(ldr >> s) & 255 for s in (0, 8, 16, 24)], -1).astype(np.uint8)
    got = np.asarray(Image.open(out).convert("RGBA"))
    assert got.shape == ref.shape
    assert np.abs(got.astype(np.int32) - ref.astype(np.int32)).max() <= 1     # double vs float camera placement: <= 1 LSB


def test_get_image_matches_read_rgba32f():
    """b2rt_get_image (CudaRenderer::getImage: renderer-owned page-locked buffer) returns the same frame as the copying
    b2rt_read_rgba32f, survives a resize, and stays valid until the next call."""
    sc = Scene.load(scene_path("CBspheres_lambertian"))
    pt = b2rt.PathTracer(ns_aa=4, max_ray_depth=3, ns_area_light=1, seed=5)
    pt.set_scene(sc)
    for (w, h) in ((96, 72), (160, 120), (64, 48)):
        pt.set_camera(place_camera(sc, w, h)); pt.set_frame_size(w, h)
        pt.render()
        a = pt.rgba32f()
        v = pt.image()
        assert v.shape == (h, w, 4) and np.array_equal(v, a)
        assert np.array_equal(pt.hdr(), a[..., :3])
        assert np.array_equal(v, a)          # the view is not disturbed by other read-backs
    pt.close()


def test_stop_leaves_an_unbiased_partial_frame():
    """b2rt_stop (PathTracer::stop, src/pathtracer.cpp:116-139): cancelled waves must not touch the sums or the
    per-pixel sample count; what was accumulated is exactly the complete waves, so the partial frame equals a frame of
    that many samples."""
    sc = Scene.load(scene_path("CBbunny"))
    w, h, spp = 128, 96, 48
    cam = place_camera(sc, w, h)
    pt = b2rt.PathTracer(ns_aa=spp, max_ray_depth=4, ns_area_light=1, seed=9, max_wave_paths=w * h)   # one sample per wave
    pt.set_scene(sc); pt.set_camera(cam); pt.set_frame_size(w, h)
    assert pt.start_raytracing()
    pt.stop()
    acc = pt.accum_tensor().cpu().numpy().reshape(h, w, 4)
    k = acc[..., 3]
    assert np.all(k == k.flat[0]) and 0 <= k.flat[0] <= spp            # whole waves only, the same count for every pixel
    done = int(k.flat[0])
    st = pt.stats()
    assert st["rays_camera"] == done * w * h
    if done:
        part = pt.hdr()
        ref = b2rt.PathTracer(ns_aa=done, max_ray_depth=4, ns_area_light=1, seed=9, max_wave_paths=w * h)
        ref.set_scene(sc); ref.set_camera(cam); ref.set_frame_size(w, h); ref.render()
        np.testing.assert_array_equal(part, ref.hdr())
        ref.close()
    # a restart after stop() begins from cleared buffers (start_raytracing clears, like the reference)
    pt.set_config(ns_aa=2)
    assert pt.start_raytracing(); pt.wait()
    acc = pt.accum_tensor().cpu().numpy().reshape(h, w, 4)
    assert np.all(acc[..., 3] == 2.0)
    pt.close()


def test_renderer_splits_waves_on_queue_overflow(monkeypatch):
    """A ray-queue overflow inside the renderer must not bias the frame: the wave is dropped on the device, rendered
    again in halves by b2rt_wait, and the result matches the frame of a run that never overflowed (up to the order of
    the fp32 sample sums)."""
    sc = Scene.load(scene_path("CBbunny"))
    w, h, spp = 160, 120, 8
    cam = place_camera(sc, w, h)
    good = b2rt.PathTracer(ns_aa=spp, max_ray_depth=4, ns_area_light=1, seed=4)
    good.set_scene(sc); good.set_camera(cam); good.set_frame_size(w, h)
    good.set_profiling(counters=True)
    good.render()
    ref = good.hdr(); st_ref = good.stats(); good.close()
    assert st_ref["queue_pushes"] > 40000
    monkeypatch.setenv("B2RT_DEBUG_PAIR_CAP", "20000")                  # far fewer pairs than one wave pushes
    pt = b2rt.PathTracer(ns_aa=spp, max_ray_depth=4, ns_area_light=1, seed=4)
    pt.set_scene(sc); pt.set_camera(cam); pt.set_frame_size(w, h); pt.render()
    img = pt.hdr(); st = pt.stats()
    acc = pt.accum_tensor().cpu().numpy().reshape(h, w, 4)
    assert np.all(acc[..., 3] == spp)                                   # every sample accumulated exactly once
    assert st["rays_camera"] == spp * w * h
    np.testing.assert_allclose(img, ref, rtol=2e-6, atol=1e-7)
    pt.close()


def test_renderer_grows_queues_on_overflow(monkeypatch):
    """A scene that pushes more rays per level than the queues were sized for: b2rt_wait enlarges the queues (kept for
    later frames) and renders the wave again whole; the frame is the one a renderer with large queues produces, bit for
    bit (one wave, so the order of the sample sums does not change)."""
    soup = random_soup(120000, size=0.05)
    sc = Scene(soup.tri_verts, materials=[dict(kind=0, albedo=(0.7, 0.6, 0.5))],
               lights=[dict(kind=1, radiance=(3.0, 3.0, 3.0), position=(0.5, 0.6, 2.0))], cam_dir=(0, 0, 1))
    w, h = 640, 480
    cam = place_camera(sc, w, h)
    cfg = dict(ns_aa=4, max_ray_depth=2, ns_area_light=1, seed=4, treelet_bytes=8192)
    monkeypatch.setenv("B2RT_RENDER_SLICE", "0")           # no distance slices: every ray is queued at every subtree it overlaps
    good = b2rt.PathTracer(**cfg)
    good.set_scene(sc); good.set_camera(cam); good.set_frame_size(w, h); good.render()
    ref = good.hdr(); st_ref = good.stats(); good.close()
    assert st_ref["queues_grown"] == 0 and st_ref["waves_retried"] == 0
    monkeypatch.setenv("B2RT_PAIR_FACTOR", "1")            # queues for one push per ray and level
    pt = b2rt.PathTracer(**cfg)
    pt.set_scene(sc); pt.set_camera(cam); pt.set_frame_size(w, h); pt.render()
    st = pt.stats()
    assert st["queues_grown"] >= 1 and st["waves_retried"] >= 1, st
    assert np.array_equal(pt.hdr(), ref)
    pt.render()                                            # the next frame runs on the enlarged queues
    st2 = pt.stats()
    assert st2["queues_grown"] == 0 and st2["waves_retried"] == 0
    pt.close()


def test_image_writers_and_one_rank_reduce(tmp_path):
    """b2rt_write_png / b2rt_write_exr store exactly what b2rt_read_ldr / b2rt_read_hdr return (top row first), and a
    one-rank NCCL reduce through the C ABI (b2rt_comm_* + b2rt_reduce_accum) leaves the accumulation buffer as it was."""
    from PIL import Image
    import sys
    sys.path.insert(0, os.path.dirname(__file__))
    from test_host import _read_exr_scanlines
    sc = Scene.load(scene_path("CBspheres_lambertian"))
    pt = b2rt.PathTracer(ns_aa=4, max_ray_depth=3, ns_area_light=1, seed=8)
    pt.set_scene(sc); pt.set_camera(place_camera(sc, 80, 60)); pt.set_frame_size(80, 60); pt.render()
    hdr, ldr = pt.hdr(), pt.ldr()
    pt.save_image(str(tmp_path / "a.png")); pt.save_exr(str(tmp_path / "a.exr"))
    got = np.asarray(Image.open(tmp_path / "a.png").convert("RGBA"))
    assert np.array_equal(got, np.stack([(ldr[::-1] >> s) & 255 for s in (0, 8, 16, 24)], -1).astype(np.uint8))
    assert np.array_equal(_read_exr_scanlines(str(tmp_path / "a.exr")), hdr[::-1])
    if b2rt.Comm.version() == 0:
        pytest.skip("no NCCL library on this box")
    before = pt.accum_tensor().clone()
    comm = b2rt.Comm(1, 0, b2rt.Comm.unique_id(), device=0)
    pt.reduce_accum(comm, root=0)
    import torch
    torch.cuda.synchronize()
    assert torch.equal(pt.accum_tensor(), before)
    assert np.array_equal(pt.hdr(), hdr)
    comm.close(); pt.close()


def test_cuda_renderer_set_viewpoint_origin_look_at():
    """CudaRenderer::setViewpoint(origin, lookAt) (src/cudaRenderer.cu:1845-1870): the reference's camera basis and
    frustum; the frame equals the oracle's for the same b2rt_camera and accumulation restarts."""
    sc = Scene.load(scene_path("CBcoil"))
    w = h = 64
    r = b2rt.CudaRenderer(samples_per_frame=2, max_ray_depth=3, ns_area_light=2, median_threshold=0, seed=2)
    r.allocOutputImage(w, h); r.loadScene(sc); r.setup()
    r.render(); r.render()
    o, L = np.array([0.0, 0.75, 3.0], np.float32), np.array([0.0, 0.0, -1.0], np.float32)   # the reference's CBcoil view
    r.setViewpoint(o, L)
    assert r.frames == 0
    r.render()
    img = r.getImage()[..., :3].copy()
    cam = b2rt.camera_look_at(o, L)
    ref = orc.OracleScene(sc, 4).render(cam, Config(ns_aa=2, max_ray_depth=3, ns_area_light=2, seed=2), w, h)
    assert np.sqrt(np.mean((img - ref) ** 2)) <= 1e-6
    assert img.max() > 0.05                          # the box is in view


def test_cpp_example_multi_gpu_matches_single(tmp_path):
    """examples/render_scene -g 2 (two handles in one process, samples dealt round-robin, b2rt_comm_create_all +
    b2rt_reduce_accum_all) writes the same HDR frame as -g 1 up to the order of the fp32 sample sums."""
    import subprocess
    import sys
    from conftest import ROOT
    sys.path.insert(0, os.path.dirname(__file__))
    from test_host import _read_exr_scanlines
    if b2rt.device_count() < 2:
        pytest.skip("needs two GPUs")
    exe = os.path.join(ROOT, "cuda-raytracer_b200", "examples", "render_scene")
    imgs = []
    for g in (1, 2):
        exr = tmp_path / f"g{g}.exr"
        r = subprocess.run([exe, "-s", "8", "-m", "4", "-l", "1", "-r", "128x96", "-g", str(g), "-w", str(tmp_path / f"g{g}.png"),
                            "-x", str(exr), scene_path("CBbunny")], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        imgs.append(_read_exr_scanlines(str(exr)))
    np.testing.assert_allclose(imgs[1], imgs[0], rtol=2e-6, atol=1e-7)
    assert imgs[0].max() > 0.1


# ---- parity pinned to the reference's own code -----------------------------------------------------------
def _reference_primary_dump(tmp_path, scene="CBbunny", size=512):
    """Run the reference's unmodified CUDA renderer (oracle/_ref/ref_cuda_render, built by oracle/build_ref.sh) for one
    frame and return its camera rays, its own traversal's result for them and its loader's triangles."""
    import subprocess
    from conftest import ROOT
    exe = os.path.join(ROOT, "oracle", "_ref", "ref_cuda_render")
    dae = os.path.join(ROOT, "oracle", "_ref", "media", scene + ".dae")
    if not (os.path.exists(exe) and os.path.exists(dae)):
        pytest.skip("oracle/_ref/ref_cuda_render not built (needs the reference checkout at build time)")
    out = tmp_path / "ref_primary.bin"
    r = subprocess.run([exe, "--dump-primary", str(out), dae, str(size)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and out.exists(), r.stderr[-2000:]
    raw = np.fromfile(out, np.uint32, 4)
    assert raw[0] == 0x31504652
    n, nt = int(raw[1]), int(raw[2])
    rays = np.fromfile(out, np.float32, n * 8, offset=16).reshape(n, 8)
    tris = np.fromfile(out, np.float32, nt * 24, offset=16 + n * 32).reshape(nt, 24)
    return rays, tris


def test_closest_hit_matches_the_reference_cuda_traversal(tmp_path):
    """The reference's OWN traversal (kernelRayIntersectSingle/Level + kernelMergeIntersections, src/cudaRenderer.cu:
    846-1297, 515-540, compiled unmodified for sm_100a) and b2rt_bvh_intersect on the SAME camera rays (the reference's
    view: eye (0, 0.75, -3), a quarter of the rays pass beside the box): the same rays hit, and the hit distances agree.
    The reference never records a primitive id (CuIntersection, src/cudaRenderer.h:155-171) and its triangle test is a
    different fp32 expression (plane + edge signs, :217-270), so the bar is on t: |dt| <= 1e-5 * max(1, t) on >= 99.99 % of
    the rays both report a hit for (observed on B200: 100 % within 1e-6); the histogram goes to gpurun_out/ref_pin_r02.json."""
    import json
    from conftest import ROOT
    rays, _ = _reference_primary_dump(tmp_path)
    o, d, t_ref, valid = rays[:, 0:3].copy(), rays[:, 3:6].copy(), rays[:, 6], rays[:, 7] > 0
    assert valid.mean() > 0.5 and np.all(np.isfinite(d))
    sc = Scene.load(scene_path("CBbunny"))
    bvh = b2rt.BVHAccel(sc)
    t, prim = bvh.intersect(o, d)
    hit = prim != 0xFFFFFFFF
    both = valid & hit
    err = np.abs(t[both] - t_ref[both]) / np.maximum(1.0, t[both])
    hist = {f"<= {b:g}": int((err <= b).sum()) for b in (0.0, 1e-7, 1e-6, 1e-5, 1e-4, 1e-3, 1e-2)}
    rec = dict(rays=int(len(rays)), reference_hits=int(valid.sum()), b2rt_hits=int(hit.sum()), both=int(both.sum()),
               only_reference=int((valid & ~hit).sum()), only_b2rt=int((~valid & hit).sum()), rel_err_hist=hist,
               max_rel_err=float(err.max()), frac_within_1e5=float((err <= 1e-5).mean()))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(rec, open(os.path.join(ROOT, "gpurun_out", "ref_pin_r02.json"), "w"), indent=1)
    assert rec["frac_within_1e5"] >= 0.9999, rec
    assert rec["only_reference"] <= 1e-4 * len(rays), rec            # a hit the reference finds and we do not would be a missed hit
    assert rec["only_b2rt"] <= 1e-3 * len(rays), rec                 # silhouette rays the reference's edge tests reject
    bvh.close()


def test_loader_matches_the_reference_loader(tmp_path):
    """The triangles the reference's own loader hands its renderer (ColladaParser::load -> DynamicScene::Mesh ->
    StaticScene::Mesh -> CuTriangle, src/cudaRenderer.cu:1679-1792, host vector CudaRenderer::triangles) against the
    scene b2rt loads for the same file: the same triangles (the reference stores each one rotated: a, b, c = p3, p1, p2),
    positions to one float ulp, per-vertex shading normals, material class.  Stated divergence: for vertices on a mesh
    BOUNDARY the reference's Vertex::normal() walks h->next()->twin() (src/halfEdgeMesh.h:625-632), which leaves the
    vertex; on this scene that only negates the normals of the 12 planar wall / light triangles, which the integrator
    cancels (shading normals are flipped towards the ray)."""
    from scipy.spatial import cKDTree
    _, tris = _reference_primary_dump(tmp_path, size=64)
    sc = Scene.load(scene_path("CBbunny"))
    nt = len(tris)
    assert nt == sc.n_tris
    ov, on = sc.tri_verts.reshape(-1, 3, 3).astype(np.float64), sc.tri_normals.reshape(-1, 3, 3).astype(np.float64)
    rv, rn = tris[:, :9].reshape(-1, 3, 3).astype(np.float64), tris[:, 9:18].reshape(-1, 3, 3).astype(np.float64)
    dist, idx = cKDTree(ov.mean(1)).query(rv.mean(1))
    assert dist.max() <= 1e-6 and len(np.unique(idx)) == nt                      # a bijection between the two triangle sets
    ours_v, ours_n = np.roll(ov[idx], -2, axis=1), np.roll(on[idx], -2, axis=1)  # the reference's vertex rotation
    assert np.abs(ours_v - rv).max() <= 2e-7
    unit = lambda v: v / np.maximum(np.linalg.norm(v, axis=-1, keepdims=True), 1e-30)
    dn = np.abs(unit(ours_n) - unit(rn)).reshape(nt, -1).max(1)
    flipped = np.abs(unit(ours_n) + unit(rn)).reshape(nt, -1).max(1) <= 1e-4
    assert np.all((dn <= 1e-4) | flipped)
    assert flipped.sum() <= 12 and (dn <= 1e-4).mean() >= 0.999
    # material class: the reference maps emitters and diffuse surfaces to fn 0 and every delta BSDF to fn 1 (:1694-1723)
    kinds = np.array([m["kind"] for m in sc.materials])[sc.tri_material]
    assert np.array_equal(np.isin(kinds, (1, 2, 4)).astype(np.int64)[idx], tris[:, 18].astype(np.int64))


def test_dae_to_frame_through_the_c_loader():
    """The whole drop-in path north_star names: .dae -> b2rt_load_dae (C++ loader) -> b2rt_set_scene -> frame, against the
    oracle rendering the scene the loader returned.  tests/golden/mini_scene.dae has triangles, a quad, an analytic
    glass sphere, a mirror block, an area light under a transformed node and a second light."""
    from conftest import ROOT
    for path, size in ((os.path.join(ROOT, "tests", "golden", "mini_scene.dae"), (72, 54)),
                       (os.path.join(ROOT, "oracle", "_ref", "media", "CBcoil.dae"), (64, 48))):
        if not os.path.exists(path):
            continue                                   # oracle/_ref/media exists only where the reference was present at build time
        sc = b2rt.load_dae(path)
        w, h = size
        cam = place_camera(sc, w, h)
        pt = b2rt.PathTracer(ns_aa=4, max_ray_depth=5, ns_area_light=2, seed=12)
        pt.set_scene(sc); pt.set_camera(cam); pt.set_frame_size(w, h); pt.render()
        img = pt.hdr()
        ref = orc.OracleScene(sc, 4).render(cam, Config(ns_aa=4, max_ray_depth=5, ns_area_light=2, seed=12), w, h)
        assert np.sqrt(np.mean((img - ref) ** 2)) <= 1e-6 and np.abs(img - ref).max() <= 1e-5, path
        assert img.max() > 0.01
        pt.close()


def test_ten_million_triangle_soup_against_brute_force():
    """BASELINE configs[4] at full size: 10 M triangles, closest hit of 1024 rays (coherent and incoherent) against the
    oracle's EXHAUSTIVE argmin over all primitives (no oracle BVH involved), on the host-built and the device-built BVH.
    (2 x 512 rays: the exhaustive search is 10^10 ray-triangle tests on the host.)"""
    sc = random_soup(10_000_000)
    rng = np.random.default_rng(77)
    n = 512
    # coherent: from a sphere of radius 2 about the cube centre towards points in the cube; incoherent: random directions
    th, ph = np.arccos(rng.uniform(-1, 1, n)), rng.uniform(0, 2 * np.pi, n)
    o1 = (np.array([.5, .5, .5]) + 2 * np.stack([np.sin(th) * np.cos(ph), np.sin(th) * np.sin(ph), np.cos(th)], 1)).astype(np.float32)
    d1 = rng.random((n, 3), dtype=np.float32) - o1
    d1 /= np.linalg.norm(d1, axis=1, keepdims=True)
    o2 = rng.random((n, 3), dtype=np.float32)
    d2 = rng.standard_normal((n, 3)).astype(np.float32)
    d2 /= np.linalg.norm(d2, axis=1, keepdims=True)
    o, d = np.concatenate([o1, o2]), np.concatenate([d1, d2]).astype(np.float32)
    t_ref, p_ref = orc.OracleScene(sc, None).intersect(o, d, mode="brute")
    assert (p_ref != 0xFFFFFFFF).mean() > 0.5
    for builder in ("gpu", "host"):
        bvh = b2rt.BVHAccel(sc, builder=builder)
        t, p = bvh.intersect(o, d)
        assert np.array_equal(p, p_ref), (builder, int((p != p_ref).sum()))
        assert np.array_equal(t, t_ref), builder
        bvh.close()


def _env_gradient(w=32, h=16):
    """A smooth HDR environment with a bright patch (so the bilinear look-up and its wrap / pole clamps all matter)."""
    y, x = np.mgrid[0:h, 0:w].astype(np.float32)
    env = np.stack([0.2 + 0.8 * x / w, 0.3 + 0.5 * y / h, 0.6 - 0.4 * x / w], -1).astype(np.float32)
    env[2:5, 3:8] += np.float32(6.0)
    return env


@pytest.mark.parametrize("scene_name", ["CBspheres_lambertian", "CBcoil"])
def test_environment_light_matches_oracle(scene_name):
    """EnvironmentLight (the `envmap` argument of PathTracer::PathTracer, src/pathtracer.h:57-60): escaping rays read the
    map, every diffuse interaction samples it as one more light.  The look-up's arctangent is the shared polynomial, so
    the frame is expected bit-identical to the oracle's; removing the map restores the plain frame."""
    sc = Scene.load(scene_path(scene_name))
    w, h = 80, 60
    cam = place_camera(sc, w, h)
    env = _env_gradient()
    cfg = dict(ns_aa=4, max_ray_depth=4, ns_area_light=1, seed=21)
    pt = b2rt.PathTracer(envmap=env, **cfg)
    pt.set_scene(sc); pt.set_camera(cam); pt.set_frame_size(w, h); pt.render()
    img = pt.hdr()
    o = orc.OracleScene(sc, 4); o.set_envmap(env)
    ref = o.render(cam, Config(**cfg), w, h)
    assert np.sqrt(np.mean((img - ref) ** 2)) <= 1e-6 and np.abs(img - ref).max() <= 1e-5
    pt.set_envmap(None); pt.render()
    o.set_envmap(None)
    plain = o.render(cam, Config(**cfg), w, h)
    assert np.sqrt(np.mean((pt.hdr() - plain) ** 2)) <= 1e-6
    assert np.abs(img - plain).max() > 0.05                 # the environment did light the scene
    pt.close()


def test_glossy_bsdf_matches_oracle():
    """GlossyBSDF(reflectance, roughness) (src/bsdf.h:143-162): Phong lobe with an integer exponent, next-event
    estimation + cosine-weighted sampling; glossy walls and a glossy sphere in the Cornell box, with and without an
    environment map."""
    from b2rt.scene import MAT_DIFFUSE, MAT_GLOSSY
    sc = Scene.load(scene_path("CBspheres_lambertian"))
    k = 0
    for m in sc.materials:
        if m["kind"] == MAT_DIFFUSE:
            m["kind"] = MAT_GLOSSY; m["roughness"] = (0.08, 0.3, 0.9)[k % 3]; k += 1
    assert k >= 3
    w, h = 72, 54
    cam = place_camera(sc, w, h)
    cfg = dict(ns_aa=4, max_ray_depth=5, ns_area_light=2, seed=31)
    for env in (None, _env_gradient()):
        pt = b2rt.PathTracer(envmap=env, **cfg)
        pt.set_scene(sc); pt.set_camera(cam); pt.set_frame_size(w, h); pt.render()
        img = pt.hdr()
        o = orc.OracleScene(sc, 4); o.set_envmap(env)
        ref = o.render(cam, Config(**cfg), w, h)
        assert np.sqrt(np.mean((img - ref) ** 2)) <= 1e-6 and np.abs(img - ref).max() <= 1e-5
        assert img.max() > 0.05
        pt.close()
