"""CPU tests of the oracle (the CPU restatement of the reference).  The reference ships no
tests or golden vectors, so the oracle is pinned by: Random123 known answers for
Philox, cv2.invert for cv::Matx33d::inv, cv2.flann for the strict radius semantics,
closed-form cases of every quirk, the committed golden vectors (tests/golden, produced by
tools/make_golden.py), and — in the dev container — the reference's own headers compiled
against a shim (oracle/ref_shim, tests/test_ref_shim_cpu.py)."""
import math
import os

import numpy as np
import pytest

from raytrace3_b200 import abi, rng
from raytrace3_b200.scene import make_scene, reference_camera, teapot_patches, teapot_transform

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_philox_known_answers(oracle):
    # Random123 kat_vectors, philox4x32-10
    kats = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
            ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
            ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
             (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for ctr, key, out in kats:
        assert oracle.philox(ctr, key) == out
        assert rng.philox4x32_10(ctr, key) == out
    for args in [(1, 2, 3, 4, 1, 5, 6), (0x5EED5EED, 7, (3 << 40) | 12345, 13, 2, 49, 31)]:
        a = oracle.draw2(*args)
        assert a == rng.draw2(*args)
        assert 0 <= a[0] < 1 and 0 <= a[1] < 1


def test_inv3_matches_cv_invert(oracle):
    cv2 = pytest.importorskip("cv2")
    r = np.random.RandomState(0)
    for i in range(2000):
        m = r.uniform(-5, 5, (3, 3)) * 10.0 ** r.randint(-3, 4)
        ok, inv = cv2.invert(m)
        assert np.array_equal(inv, oracle.inv3(m)), i
    sing = np.array([[1., 2, 3], [2, 4, 6], [0, 1, 5]])
    assert np.array_equal(oracle.inv3(sing), np.zeros((3, 3)))  # det == 0 -> zeros (Matx::inv)
    assert np.array_equal(cv2.invert(sing)[1], np.zeros((3, 3)))


def test_flann_radius_is_strict(oracle):
    """Q3: flann radiusSearch takes a squared radius and excludes dist^2 == radius."""
    cv2 = pytest.importorskip("cv2")
    pts = np.array([[0, 0, 0], [1, 1, 0], [1, 0, 0], [3, 0, 0]], dtype=np.float32)
    idx = cv2.flann_Index(pts, dict(algorithm=4))  # KDTREE_SINGLE (the reference's index type)
    cnt, ind, d = idx.radiusSearch(np.zeros((1, 3), np.float32), 2.0, 4)
    assert cnt == 2 and sorted(int(i) for i in ind[0][:cnt]) == [0, 2]  # (1,1,0) at dist^2 == 2.0 is excluded
    hp = dict(pos=pts.astype(np.float64), n=np.tile([0., 0, 1], (4, 1)), wgt=np.ones((4, 3)), r2=np.full(4, 2.0))
    dep = dict(x=np.zeros((1, 3)), n=np.array([[0., 0, 1]]), flux=np.ones((1, 3)))
    newN, tao, cand, acc = oracle.gather_brute(hp, dep, 2.0)
    assert list(newN) == [1, 0, 1, 0] and cand == 2


def test_aabb_entry_face_only(oracle):
    """Q5: AABBbox::intersect tests entry faces only; an origin inside the box misses."""
    box = [0, 0, 0, 1, 1, 1]
    assert oracle.aabb_intersect(box, [-1, .5, .5], [1, 0, 0])
    assert not oracle.aabb_intersect(box, [.5, .5, .5], [1, 0, 0])       # inside -> miss
    assert not oracle.aabb_intersect(box, [-1, .5, .5], [-1, 0, 0])     # pointing away
    assert not oracle.aabb_intersect(box, [-1, 2.5, .5], [1, 0, 0])
    assert oracle.aabb_intersect(box, [-1, -1, .5], [1, 1, 0.00001])    # dir.z below 1e-4 is skipped
    assert not oracle.aabb_intersect(box, [.5, .5, -1], [0, 0, 0.00009])  # IsZero(dir) on every axis


def test_bezier_eval_against_polynomial(oracle):
    P = teapot_patches()[5].reshape(16, 3)
    B = lambda t: np.array([(1 - t) ** 3, 3 * t * (1 - t) ** 2, 3 * t * t * (1 - t), t ** 3])
    dB = lambda t: np.array([-3 * (1 - t) ** 2, 3 * (1 - t) ** 2 - 6 * t * (1 - t), 6 * t * (1 - t) - 3 * t * t, 3 * t * t])
    for u, v in [(0.3, 0.7), (0, 0), (1, 1), (0.999, 0.001)]:
        S, Su, Sv = oracle.bezier_eval(P, u, v)
        G = P.reshape(4, 4, 3)  # P[4*i+k]: i <-> v, k <-> u (Bezier.h:85-90)
        np.testing.assert_allclose(S, np.einsum("i,k,ikc->c", B(v), B(u), G), rtol=1e-13, atol=1e-13)
        np.testing.assert_allclose(Su, np.einsum("i,k,ikc->c", B(v), dB(u), G), rtol=1e-12, atol=1e-12)
        np.testing.assert_allclose(Sv, np.einsum("i,k,ikc->c", dB(v), B(u), G), rtol=1e-12, atol=1e-12)


def _flat_patch(z=0.0):
    g = np.zeros((4, 4, 3))
    for i in range(4):
        for k in range(4):
            g[i, k] = (k / 3.0, i / 3.0, z)
    return g.reshape(16, 3)


def test_newton_quirks(oracle):
    """Q4: every accepted iterate is recorded; the minimum t wins even when negative."""
    P = _flat_patch(0.0)  # S(u,v) = (u, v, 0)
    seeds = np.full((50, 2), 0.5)
    # ray hits the plane z=0 at (0.25, 0.75) with t = 2: Newton on a linear map converges in 1 step
    out = oracle.newton(P, [0.25, 0.75, 2.0], [0, 0, -1.0], seeds)
    np.testing.assert_allclose(out, [2.0, 0.25, 0.75], atol=1e-12)
    # solution behind the origin: newton returns the negative t (the caller rejects t < 1e-4)
    out = oracle.newton(P, [0.25, 0.75, -2.0], [0, 0, -1.0], seeds)
    np.testing.assert_allclose(out, [-2.0, 0.25, 0.75], atol=1e-12)
    # outside [0,1]^2: nothing recorded -> MAX_VEC3
    out = oracle.newton(P, [1.5, 0.5, 2.0], [0, 0, -1.0], seeds)
    assert (out == 1e20).all()
    # singular Jacobian (ray parallel to the patch): inv() returns zeros, the iterate stalls
    out = oracle.newton(P, [0.25, 0.75, 2.0], [1.0, 0, 0], seeds)
    assert (out == 1e20).all()
    # ... but a stalled iterate that already sits within the 0.01 residual ball IS recorded
    out = oracle.newton(P, [0.5, 0.5, 0.005], [1.0, 0, 0], seeds)
    np.testing.assert_array_equal(out, [0.0, 0.5, 0.5])


def test_color_uv_wrap(oracle):
    t = np.zeros((4, 5, 3), np.uint8)
    t[..., 0] = np.arange(20).reshape(4, 5) * 10
    t[..., 1] = 255
    t[..., 2] = 7
    c = oracle.color_uv(t, 0.0, 0.0)
    np.testing.assert_allclose(c, [0, 1, 7 / 255.0], atol=1e-15)
    c = oracle.color_uv(t, 0.125, 0.1)  # row 0.5, col 0.5: mean of (0,0),(0,1),(1,0),(1,1)
    np.testing.assert_allclose(c[0], (0 + 10 + 50 + 60) / 4 / 255.0, atol=1e-12)
    c1 = oracle.color_uv(t, 0.875, 0.9)  # r1=3,r2 wraps to 0; c1=4,c2 wraps to 0
    np.testing.assert_allclose(c1[0], (190 + 150 + 40 + 0) / 4 / 255.0, atol=1e-12)
    np.testing.assert_allclose(oracle.color_uv(t, 1.125, -0.9), oracle.color_uv(t, 0.125, 0.1), atol=1e-12)


def test_sampling_and_optics(oracle):
    n = np.array([0.0, 0.0, 1.0])
    for u1, u2 in [(0.3, 0.6), (0.999999, 0.1), (1e-9, 0.9)]:
        d = oracle.unit_rand_refl(n, u1, u2)
        assert abs(np.linalg.norm(d) - 1) < 1e-12
        assert abs(d[2] - math.sqrt(u1)) < 1e-7  # cos(theta) = sqrt(U1) (PI is truncated -> 1e-7)
    n2 = np.array([0.6, 0.0, 0.8])
    d = oracle.unit_rand_refl(n2, 0.25, 0.4)
    assert abs(d @ n2 - 0.5) < 1e-7
    # Snell: normal incidence passes straight through; grazing from inside -> TIR mirror (Q17)
    I = np.array([0.0, 0.0, -1.0])
    np.testing.assert_allclose(oracle.refracted(I, n, 1 / 1.5), I, atol=1e-15)
    I = np.array([math.sin(1.0), 0.0, -math.cos(1.0)])
    out = oracle.refracted(I, n, 1.5)  # sin(1)*1.5 > 1 -> total internal reflection
    np.testing.assert_allclose(out, [I[0], 0, -I[2]], atol=1e-15)
    out = oracle.refracted(I, n, 1 / 1.5)
    assert abs(out[0] - math.sin(1.0) / 1.5) < 1e-12 and out[2] < 0


def test_camera_matches_python_mirror(oracle):
    cam = reference_camera(320, 240)
    oc = oracle.camera_lookat(cam.pos, cam.look, 320, 240)
    for k in ("pos", "dir", "du", "dv"):
        assert list(getattr(oc, k)) == list(getattr(cam, k)), k
    # dir=(0,0.04257,-0.99909)*1.07225, du=(1,0,0), dv=(0,0.99909,0.04257)
    np.testing.assert_allclose(cam.du, [1, 0, 0], atol=1e-15)
    np.testing.assert_allclose(cam.dv, [0, 0.99909, 0.04257], atol=1e-5)
    np.testing.assert_allclose(np.linalg.norm(cam.dir), 0.5 / math.tan(25 * 3.1415926 / 180), rtol=1e-15)
    r = oracle.camera_emit(cam.desc(), 0, 0)
    assert abs(np.linalg.norm(r[3:]) - 1) < 1e-15 and r[3] < 0 and r[4] < 0  # Q12/Q13: row 0 = bottom


def test_teapot_world_bbox():
    """world bbox after Scene.h:142-153 ~ [12,0,105.9]-[28,12.6,132]."""
    sc = make_scene("reference", textures=None)
    P = sc.objvec[8].P.reshape(-1, 3)
    np.testing.assert_allclose(P.min(0), [12, 0, 105.9], atol=0.01)
    np.testing.assert_allclose(P.max(0), [28, 12.6, 132], atol=0.01)
    T = teapot_transform()
    assert abs(T[0, 1] - 1) < 1e-12 and abs(T[1, 2] - 1) < 1e-12 and abs(T[2, 0] + 1) < 1e-12


def test_scene_statistics(oracle, params):
    """SURVEY §8(d) probe: 1.000 hit points per pixel, ~5.9 queries and ~3.7 deposits per photon."""
    sc = make_scene("cornell", textures=None)
    cam = reference_camera(64, 48)
    hp = oracle.eye_pass(sc, params, cam)
    assert hp["pos"].shape[0] == 64 * 48
    assert (hp["r2"] == 2.0).all()
    r = oracle.photon_pass(sc, params, 0, 3000, 0, hp=hp)
    c = r["counters"]
    assert 5.6 < c["queries"] / c["photons"] < 6.2
    assert 3.5 < c["deposits"] / c["photons"] < 3.9
    assert c["queries"] <= 15 * c["photons"]  # Q9
    assert r["newN"].sum() > 0


def test_kdtree_equals_brute(oracle, params):
    r = np.random.RandomState(2)
    n, m = 3000, 4000
    hp = dict(pos=r.uniform(0, 20, (n, 3)), n=np.tile([0., 1, 0], (n, 1)), wgt=r.uniform(size=(n, 3)), r2=r.uniform(.5, 2, n))
    hp["pos"][:50] = hp["pos"][0]  # duplicates
    dep = dict(x=r.uniform(0, 20, (m, 3)), n=np.tile([0., 1, 0], (m, 1)), flux=r.uniform(size=(m, 3)))
    a, ta, _, _ = oracle.gather_brute(hp, dep, 2.0)
    b, tb = oracle.gather_kdtree(hp, dep, 2.0)
    assert a.sum() > 100
    np.testing.assert_array_equal(a, b)
    np.testing.assert_allclose(ta, tb, rtol=1e-12)


def test_update_reference_is_noop_and_progressive_truncates(oracle):
    newN = np.array([5, 0, 3], np.int32)
    N = np.zeros(3, np.int32)
    r2 = np.full(3, 2.0)
    tao = np.ones((3, 3))
    oracle.update(abi.SPPM_UPDATE_REFERENCE, 0.7, newN, N, r2, tao)
    assert list(newN) == [5, 0, 3] and list(N) == [0, 0, 0] and (r2 == 2.0).all()  # Q1
    oracle.update(abi.SPPM_UPDATE_PROGRESSIVE, 0.7, newN, N, r2, tao)
    assert list(N) == [3, 0, 2] and list(newN) == [0, 0, 0]  # int(0.7*5)=3, int(0.7*3)=2
    np.testing.assert_allclose(r2, [2 * 0.7, 2.0, 2 * 0.7])
    newN[:] = [4, 1, 0]
    oracle.update(abi.SPPM_UPDATE_REFERENCE, 0.7, newN, N, r2, tao)  # now N>0 where it fired
    assert list(N) == [int(3 + 0.7 * 4), 0, 2]
    np.testing.assert_allclose(r2[0], 2 * 0.7 * (3 + 0.7 * 4) / 7)


def test_resolve_and_tonemap(oracle):
    pix = np.array([[0, 0], [0, 0], [1, 2]], np.int32)
    r2 = np.array([2.0, 1.0, 2.0])
    tao = np.array([[1., 2, 3], [1, 1, 1], [4, 4, 4]])
    film = oracle.resolve(pix, r2, tao, 2, 10, 3, 2)
    s = 1.0 / 3.1415926
    np.testing.assert_allclose(film[0, 0], [s / 40 + s / 20, 2 * s / 40 + s / 20, 3 * s / 40 + s / 20], rtol=1e-14)
    np.testing.assert_allclose(film[1, 2], 4 * s / 40, rtol=1e-14)
    img = oracle.tonemap(film, 1.0)
    assert img.shape == (2, 3, 3)
    assert img[0, 2, 0] == int((1 - math.exp(-film[1, 2, 0])) ** (1 / 2.2) * 255 + .5)  # rows flipped
    assert img[1, 1].max() == 0


@pytest.mark.parametrize("name", ["newton", "near_collide", "eye_cornell", "photon_cornell"])
def test_golden_vectors(oracle, params, name):
    """The committed golden vectors (tools/make_golden.py) still come out of the oracle."""
    g = np.load(os.path.join(GOLD, name + ".npz"))
    if name == "newton":
        sc = make_scene("reference", textures=None)
        out = oracle.bezier_patch_batch(sc, params, g["rays"], g["patch"], g["seeds"])
        np.testing.assert_array_equal(out, g["tuv"])
    elif name == "near_collide":
        sc = make_scene("reference", textures=None)
        out = oracle.intersect_batch(sc, params, g["rays"], g["ids"], pass_id=2, node=5)
        for k in ("t", "obj", "patch", "u", "v", "pos"):
            np.testing.assert_array_equal(out[k], g[k])
    elif name == "eye_cornell":
        sc = make_scene("cornell", textures=None)
        hp = oracle.eye_pass(sc, params, reference_camera(32, 24), pass_id=0)
        for k in ("pos", "n", "wgt", "pix"):
            np.testing.assert_array_equal(hp[k], g[k])
    else:
        sc = make_scene("cornell", textures=None)
        hp = {k: g["hp_" + k] for k in ("pos", "n", "wgt", "r2")}
        r = oracle.photon_pass(sc, params, 0, 500, 3, hp=hp, keep_deposits=True)
        np.testing.assert_array_equal(r["newN"], g["newN"])
        np.testing.assert_allclose(r["tao"], g["tao"], rtol=1e-12)
        np.testing.assert_allclose(r["dep"]["x"], g["dep_x"], rtol=0, atol=1e-9)


def test_oracle_is_clean_under_asan_ubsan():
    """the same oracle source built with -fsanitize=address,undefined (`make -C oracle asan`)
    passes the golden-vector tests without a sanitizer report (out-of-bounds reads would silently
    bend a checker that everything else is held to)."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    asan = subprocess.run([cxx, "-print-file-name=libasan.so"], capture_output=True, text=True).stdout.strip()
    if not os.path.isabs(asan) or not os.path.exists(asan):
        pytest.skip("libasan.so not installed")
    r = subprocess.run(["make", "-C", os.path.join(root, "oracle"), "asan"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    env = dict(os.environ, LD_PRELOAD=asan, ASAN_OPTIONS="detect_leaks=0:halt_on_error=1", UBSAN_OPTIONS="print_stacktrace=1",
               SPPM_ORACLE_LIB=os.path.join(root, "oracle", "_build", "liboracle_asan.so"))
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(root, "tests", "test_ref_golden_cpu.py"), "-q", "-s", "-x",
                        "-p", "no:cacheprovider"], capture_output=True, text=True, env=env, cwd=root, timeout=600)
    out = r.stdout + r.stderr
    assert r.returncode == 0, out[-3000:]
    assert "runtime error" not in out and "AddressSanitizer" not in out, out[-3000:]
