"""Pins the oracle against THE REFERENCE ITSELF: oracle/_ref/libref.so is the reference's own
headers (/root/reference/raytracer/*.h) compiled against a thin OpenCV/FLANN shim
(oracle/ref_shim), with the random numbers injected in the order the reference consumes them.
Every comparison below is bit-exact: the oracle restates the same arithmetic and both run on
the host libm.  Skipped where the library was not built (it needs /root/reference)."""
import math

import numpy as np
import pytest

from conftest import custom_scene as _custom_scene, quirks_scene as _quirks_scene, rays_at_patches, rays_at_teapot, rays_in_room
from raytrace3_b200 import abi
from raytrace3_b200.scene import Material, PlaneObj, Scene, SphereObj, SpotLight, make_scene, reference_camera, teapot_patches

ref_mod = pytest.importorskip("oracle.ref")
pytestmark = pytest.mark.skipif(not ref_mod.available(), reason="oracle/_ref/libref.so not built (needs /root/reference)")


@pytest.fixture(scope="module")
def ref():
    return ref_mod.Reference()


def test_is_the_real_isZero(ref):
    """Sanity of the build: IsZero (Vec3.h:23) must use the double abs, not int abs."""
    assert ref.aabb_intersect([0, 0, 0, 1, 1, 1], [-1, .5, .5], [0.5, 0, 0])       # |0.5| is not "zero"
    assert not ref.aabb_intersect([0, 0, 0, 1, 1, 1], [-1, .5, .5], [0.00009, 0, 0])  # below 1e-4 is


def test_bezier_eval_and_aabb(oracle, ref):
    rng = np.random.RandomState(0)
    P = teapot_patches()
    for k in range(300):
        p = P[rng.randint(32)] * rng.uniform(0.5, 5)
        u, v = rng.uniform(-0.2, 1.2, 2)
        for a, b in zip(oracle.bezier_eval(p, u, v), ref.bezier_eval(p, u, v)):
            assert np.array_equal(a, b)
    for k in range(3000):
        lo = rng.uniform(-5, 5, 3)
        box = np.concatenate([lo, lo + rng.uniform(0.1, 6, 3)])
        org = rng.uniform(-12, 12, 3)
        d = rng.normal(size=3)
        if k % 5 == 0:
            d[rng.randint(3)] = rng.choice([0.0, 5e-5, -5e-5, 1e-4])  # IsZero(dir) edge
        if k % 7 == 0:
            org = lo + rng.uniform(0, 0.1, 3)  # origin inside / on the box (Q5)
        d /= np.linalg.norm(d)
        assert oracle.aabb_intersect(box, org, d) == ref.aabb_intersect(box, org, d)


def test_newton_bit_identical(oracle, ref):
    """Bezier3::newton (Bezier.h:112-159) incl. cv::Matx33d::inv, all 50x10 iterates, min over
    recorded iterates (Q4)."""
    sc = make_scene("reference", textures=None)
    rng = np.random.RandomState(3)
    n = 250
    patch = rng.randint(0, 32, n)
    rays = rays_at_patches(patch, 5, sc)
    rays[:40] = rays_at_teapot(40, 6, sc)
    P = sc.objvec[8].P
    hits = 0
    for i in range(n):
        seeds = rng.uniform(size=(50, 2))
        a = oracle.newton(P[patch[i]], rays[i, :3], rays[i, 3:], seeds)
        b = ref.newton(P[patch[i]], rays[i, :3], rays[i, 3:], seeds)
        assert np.array_equal(a, b), (i, a, b)
        hits += a[0] < 1e19
    assert 100 < hits < n
    # flat patch: singular Jacobian (inv() -> zeros) and a negative-t solution
    g = np.zeros((4, 4, 3))
    for i in range(4):
        for k in range(4):
            g[i, k] = (k / 3.0, i / 3.0, 0.0)
    seeds = np.full((50, 2), 0.5)
    for org, d in [([0.25, 0.75, 2.0], [0, 0, -1.0]), ([0.25, 0.75, -2.0], [0, 0, -1.0]), ([0.25, 0.75, 2.0], [1.0, 0, 0]),
                   ([0.5, 0.5, 0.005], [1.0, 0, 0])]:
        assert np.array_equal(oracle.newton(g.reshape(16, 3), org, d, seeds), ref.newton(g.reshape(16, 3), org, d, seeds))


def test_sampling_optics_camera(oracle, ref):
    rng = np.random.RandomState(4)
    for k in range(500):
        n = rng.normal(size=3)
        n /= np.linalg.norm(n)
        if k % 10 == 0:
            n = np.array([0.0, 0.0, rng.choice([-1.0, 1.0])])  # GetAnormal special case (Vec3.h:87)
        u1, u2 = rng.uniform(size=2)
        if k % 17 == 0:
            u1 = 1.0 - 1e-9  # theta below 1e-4: rotated() returns *this (Vec3.h:101)
        assert np.array_equal(oracle.unit_rand_refl(n, u1, u2), ref.unit_rand_refl(n, u1, u2))
        I = rng.normal(size=3)
        I /= np.linalg.norm(I)
        N = n if I @ n < 0 else -n
        for eta in (1.5, 1 / 1.5):
            assert np.array_equal(oracle.refracted(I, N, eta), ref.refracted(I, N, eta))
    for w, h in [(320, 240), (1024, 768), (64, 48)]:
        pos = np.array([50., 35., 230.]) + rng.uniform(-1, 1, 3) * (w == 64)
        look = pos + np.array([0, 0.042612, -1.])
        oc = oracle.camera_lookat(pos, look, w, h)
        rc, rray = ref.camera(pos, look, w, h, y=h // 3, x=w // 5)
        for f in ("pos", "dir", "du", "dv"):
            assert list(getattr(oc, f)) == list(getattr(rc, f)), f
        assert np.array_equal(oracle.camera_emit(oc, h // 3, w // 5), rray)


def test_texture_update_tonemap(oracle, ref):
    rng = np.random.RandomState(5)
    tex = rng.randint(0, 256, size=(13, 7, 3)).astype(np.uint8)
    for k in range(400):
        u, v = rng.uniform(-3, 3, 2)
        if k % 9 == 0:
            u, v = float(rng.randint(-2, 3)), rng.choice([0.0, -1e-20, 1 - 1e-16])
        assert np.array_equal(oracle.color_uv(tex, u, v), ref.color_uv(tex, u, v))
    n = 400
    newN = rng.randint(0, 40, n).astype(np.int32)
    N = rng.randint(0, 30, n).astype(np.int32)
    N[:100] = 0  # the as-shipped case (Q1)
    r2 = rng.uniform(0.1, 2, n)
    tao = rng.uniform(0, 50, (n, 3))
    a = [x.copy() for x in (newN, N, r2, tao)]
    b = [x.copy() for x in (newN, N, r2, tao)]
    oracle.update(abi.SPPM_UPDATE_REFERENCE, 0.7, *a)
    ref.update(*b)
    for x, y in zip(a, b):
        assert np.array_equal(x, y)
    film = rng.uniform(0, 3, (4, 5, 3))
    img = oracle.tonemap(film, 4.0)
    for i in range(4):
        for j in range(5):
            for k in range(3):
                assert img[3 - i, j, k] == ref.to_int(film[i, j, k] / 4.0)
    pix = np.array([[0, 0]], np.int32)
    f = oracle.resolve(pix, np.array([1.7]), np.array([[3.0, 4.0, 5.5]]), 7, 100, 1, 1)
    assert np.array_equal(f[0, 0], ref.resolve_one([3.0, 4.0, 5.5], 1.7, 7, 100))


def _scene(name):
    if name == "custom":
        return _custom_scene()
    if name == "quirks":
        return _quirks_scene()
    return make_scene(name, textures="synthetic" if name == "reference" else None)


@pytest.mark.parametrize("name", ["cornell", "reference", "stress", "custom", "quirks"])
def test_near_collide_bit_identical(oracle, ref, params, name):
    """Scene::NearCollide (Scene.h:165-182) through the reference's own PlaneObj / SphereObj /
    Bezier3Obj (.bpt loader included), textures included."""
    sc = _scene(name)
    h = ref.scene(sc)
    n_room, n_tp = (250, 0) if name in ("cornell", "custom") else (120, 130)
    rays = rays_in_room(n_room, 7)
    if n_tp:
        rays = np.concatenate([rays, rays_at_teapot(n_tp, 8, sc)])
    bez = 0
    for i in range(rays.shape[0]):
        oracle.log_begin()
        a = oracle.intersect_batch(sc, params, rays[i:i + 1], np.array([i], np.uint64), pass_id=1, node=2)[0]
        vals, kinds = oracle.log_end()
        b = ref.near_collide(h, rays[i], vals, kinds)
        assert a["obj"] == b["obj"] and a["inside"] == b["inside"], i
        for f in ("t", "pos", "n", "color"):
            assert np.array_equal(a[f], b[f]), (i, f, a[f], b[f])
        bez += a["patch"] >= 0
    if n_tp:
        assert bez > 20


@pytest.mark.parametrize("name", ["cornell", "reference", "custom", "quirks"])
def test_photon_trace_bit_identical(oracle, ref, params, name):
    """Light::emit + PhotonTrace + radiusSearch/accumulate (Light.h:8-13, Raytracer.h:117-209) for
    single photons: same emission, same bounces, same newN and tao on every hit point."""
    sc = _scene(name)
    h = ref.scene(sc)
    hp = oracle.eye_pass(sc, params, reference_camera(96, 72), 0)
    total = 0
    for j in range(150):
        oracle.log_begin()
        r = oracle.photon_pass(sc, params, j, 1, 5, hp=hp, use_kdtree=False)  # one photon per light, lights in order
        vals, kinds = oracle.log_end()
        newN, tao = ref.photon(h, len(sc.plvec), hp, vals, kinds)
        assert np.array_equal(r["newN"], newN), j
        assert np.array_equal(r["tao"], tao), j
        total += int(newN.sum())
    assert total > 10  # photons really landed on hit points


def test_view_trace_bit_identical(oracle, ref, params):
    """ViewTrace (Raytracer.h:296-357): mirror / glass chains, branching materials, the teapot."""
    for name in ("reference", "custom", "quirks"):
        sc = _scene(name)
        h = ref.scene(sc)
        w, hh = 40, 30
        cam = reference_camera(w, hh)
        count = 0
        for i in range(0, hh, 2):
            for j in range(0, w, 3):
                oracle.log_begin()
                a = oracle.view_trace_pixel(sc, params, cam, i, j)
                vals, kinds = oracle.log_end()
                b = ref.view_trace(h, cam.pos, cam.look, w, hh, i, j, vals, kinds)
                assert a["pos"].shape == b["pos"].shape, (i, j)
                for f in ("pos", "n", "wgt", "r2", "pix"):
                    assert np.array_equal(a[f], b[f]), (name, i, j, f)
                count += a["pos"].shape[0]
        if name == "quirks":   # Q8: eye paths ending on the objects with a zero diffuse channel make no hit point
            assert 0 < count < (hh // 2) * (w // 3 + 1)
        else:
            assert count >= (hh // 2) * (w // 3 + 1) * 0.9
        if name == "custom":
            assert count > (hh // 2) * (w // 3 + 1)  # branching produced extra hit points


def test_flann_shim_tree_equals_scan(ref):
    """The kd-tree the FLANN shim builds (so that the reference's own PhotonMap can be timed with an
    index of the right complexity) returns exactly the exhaustive scan's (index, distance) lists:
    random clouds with points on shared planes and duplicates, queries on and next to data points."""
    if not ref.can_time:
        pytest.skip("libref.so predates the timing entry points")
    compared = 0
    for n, nq, side, seed in ((5, 40, 2.0, 1), (11, 60, 1.0, 2), (3000, 800, 12.0, 3), (60000, 600, 40.0, 4)):
        bad, found = ref.flann_selfcheck(n, nq, side, 2.0, seed)
        assert bad == 0, (n, bad)
        compared += found
    assert compared > 10000


def test_reference_photon_map_runs_and_is_timed(ref):
    """RayTracer::GetHitPoint + index build + PhotonMap(cnt) of the reference's own code on two replicas
    (bench.py's cpu_baseline.reference_code leg): one hit point per pixel in the reference scene,
    100 photons per round and light, stdout left usable afterwards."""
    if not ref.can_time:
        pytest.skip("libref.so predates the timing entry points")
    sc = make_scene("reference", textures=None)
    cam = reference_camera(48, 36)
    o = ref.time_photon_map(ref.scene(sc), cam.pos, cam.look, cam.w, cam.h, 7, 2)
    assert o["hit_points"] == 48 * 36 and o["rounds"] == 7 and o["photons"] == 2 * 7 * 100 * len(sc.plvec)
    assert o["map_s"] > 0 and o["setup_s"] > 0
    o = ref.time_photon_map(ref.scene(sc), cam.pos, cam.look, cam.w, cam.h, 100000, 2, target_s=0.2)
    assert 10 <= o["rounds"] < 100000 and o["map_s"] < 5.0
