"""The PRODUCT's device source on the CPU: raytrace3_b200/csrc/{device_math.cuh, scene.cuh} and the
host table builder scene_prep.h compiled by g++ through tests/device_host/cuda_host_shim.h (a warp
= 32 fibers switched at every warp collective), held to the oracle and to the golden vectors the
reference's own code produced.  These are the same source lines the sm_100a kernels are built
from, so the arithmetic, the division-free rejects, the Newton culls, the item queue and the
segmented arg-min of near_collide are fuzzed here at sizes that would cost GPU minutes; what this
cannot see is nvcc's code generation (the `-m gpu` tests hold that)."""
import os

import numpy as np
import pytest

from conftest import custom_scene, quirks_scene, rays_at_patches, rays_at_teapot, rays_in_room
from raytrace3_b200 import abi
from raytrace3_b200.scene import make_scene, synthetic_texture

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def dh():
    from device_host import DeviceHost
    return DeviceHost()


def _params(**kw):
    p = abi.default_params()
    for k, v in kw.items():
        setattr(p, k, v)
    return p


def _patches(sc):
    return np.concatenate([o.P for o in sc.objvec if o.kind == 2], axis=0).reshape(-1, 48)


def _scene(name):
    if name == "custom":
        return custom_scene()
    if name == "quirks":
        return quirks_scene()
    return make_scene(name, textures="synthetic" if name == "reference" else None)


# ---- the counter RNG and the samplers ---------------------------------------------------------
def test_draw2_is_the_oracles_philox(dh, oracle):
    rng = np.random.RandomState(1)
    for _ in range(300):
        a = [int(rng.randint(0, 2 ** 62)) for _ in range(3)] + [int(rng.randint(0, 2 ** 31)) for _ in range(4)]
        assert dh.draw2(*a) == oracle.draw2(*a)


def test_unit_rand_refl_and_refracted(dh, oracle):
    """Vec3::UnitRandRefl / Refracted (Vec3.h:57-65, 90-115, 124-134), incl. total internal reflection."""
    rng = np.random.RandomState(2)
    n = 2000
    nrm = rng.normal(size=(n, 3))
    nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
    nrm[:6] = np.array([[1, 0, 0], [0, 1, 0], [0, 0, 1], [-1, 0, 0], [0, -1, 0], [0, 0, -1]], dtype=float)  # axis branches
    u12 = rng.uniform(size=(n, 2))
    got = dh.unit_rand_refl(nrm, u12, fast=False)
    ref = np.stack([oracle.unit_rand_refl(nrm[i], u12[i, 0], u12[i, 1]) for i in range(n)])
    np.testing.assert_array_equal(got, ref)
    fast = dh.unit_rand_refl(nrm, u12, fast=True)       # SPPM_MATH_FAST: sincospi-style evaluation
    np.testing.assert_allclose(fast, ref, rtol=0, atol=1e-12)
    I = rng.normal(size=(n, 3))
    I /= np.linalg.norm(I, axis=1, keepdims=True)
    nn = np.where(rng.uniform(size=n) < 0.5, 1.5, 1 / 1.5)
    got = dh.refracted(I, nrm, nn)
    ref = np.stack([oracle.refracted(I[i], nrm[i], nn[i]) for i in range(n)])
    np.testing.assert_array_equal(got, ref)


def test_color_uv_bilinear(dh, oracle):
    """Texture::colorUV (Element.h:61-72) through the one-word-per-texel device layout."""
    tex = synthetic_texture(37, 53, 11)
    rng = np.random.RandomState(3)
    uv = rng.uniform(-0.2, 1.2, size=(3000, 2))
    uv[:8] = [[0, 0], [1, 1], [0, 1], [1, 0], [0.5, 0.5], [1 - 1e-16, 1 - 1e-16], [1e-300, 0], [0.999999, 0.000001]]
    got = dh.color_uv(tex, uv)
    ref = np.stack([oracle.color_uv(tex, u, v) for u, v in uv])
    np.testing.assert_array_equal(got, ref)


# ---- boxes, patch evaluation, Newton ----------------------------------------------------------
def test_aabb_intersect_fuzz(dh, oracle):
    """AABBbox::intersect (Bezier.h:20-51), incl. axis-parallel rays and origins inside the box."""
    rng = np.random.RandomState(4)
    box = np.array([-3.0, 1.0, 2.0, 4.0, 5.0, 2.5])
    n = 20000
    org = rng.uniform(-12, 12, size=(n, 3))
    d = rng.normal(size=(n, 3))
    tgt = box[:3] + (box[3:] - box[:3]) * rng.uniform(-0.3, 1.3, size=(n // 2, 3))
    d[:n // 2] = tgt - org[:n // 2]                      # half aimed at (or just past) the box
    d[rng.uniform(size=(n, 3)) < 0.15] = 0.0           # zero components
    d[np.abs(d).sum(1) == 0] = [0, 0, 1]
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    rays = np.concatenate([org, d], axis=1)
    got = dh.aabb_intersect(box, rays)
    ref = np.array([oracle.aabb_intersect(box, r[:3], r[3:]) for r in rays])
    np.testing.assert_array_equal(got, ref)
    assert 2000 < ref.sum() < n - 2000


def test_bezier_eval_three_modes(dh, oracle):
    """evalPatchPoint / DerivU / DerivV (Bezier.h:85-111): EXACT bit for bit; the monomial (FAST)
    and the factored Bernstein (FAST_BERNSTEIN) forms to 1e-12 of the patch's scale."""
    sc = make_scene("reference", textures=None)
    P = _patches(sc)
    hs = dh.scene(sc)
    A = hs.tables()["A"]
    rng = np.random.RandomState(5)
    uv = rng.uniform(-0.5, 1.5, size=(400, 2))   # Newton iterates leave [0,1]^2
    uv[:4] = [[0, 0], [1, 1], [0, 1], [1, 0]]
    for p in range(P.shape[0]):
        ref = np.stack([np.concatenate(oracle.bezier_eval(P[p], u, v)) for u, v in uv])
        S, Su, Sv = dh.bezier_eval(0, P[p], uv)
        np.testing.assert_array_equal(np.concatenate([S, Su, Sv], axis=1), ref)
        scale = np.abs(P[p]).max()
        for mode, coef in ((1, A[p]), (2, P[p])):
            S, Su, Sv = dh.bezier_eval(mode, coef, uv)
            np.testing.assert_allclose(np.concatenate([S, Su, Sv], axis=1), ref, rtol=0, atol=2e-12 * scale)


def test_newton_exact_is_the_reference_bit_for_bit(dh, oracle):
    """Bezier3::newton (Bezier.h:112-159): the kernels' newton_restart<EXACT> + min over restarts
    == the golden vectors of the reference's own code, and == the oracle on a larger fuzz."""
    g = np.load(os.path.join(GOLDEN, "ref_newton.npz"))
    sc = make_scene("reference", textures=None)
    P = _patches(sc)
    for p in np.unique(g["patch"]):
        sel = g["patch"] == p
        got = dh.newton(0, P[p], g["rays"][sel], g["seeds"][sel])
        np.testing.assert_array_equal(got, g["tuv"][sel])
    prm = _params()
    rng = np.random.RandomState(6)
    n = 6000
    patch = np.sort(rng.randint(0, P.shape[0], size=n)).astype(np.int32)
    rays = rays_at_patches(patch, 7, sc)
    rays[::7] = rays_at_teapot(len(rays[::7]), 8, sc)
    seeds = rng.uniform(size=(n, 20, 2))
    prm.newton_restarts = 20
    ref = oracle.bezier_patch_batch(sc, prm, rays, patch, seeds)
    for p in range(P.shape[0]):
        sel = patch == p
        np.testing.assert_array_equal(dh.newton(0, P[p], rays[sel], seeds[sel]), ref[sel])
    assert (ref[:, 0] < 1e19).sum() > 3000 and (ref[:, 0] >= 1e19).sum() > 300


@pytest.mark.parametrize("mode", [abi.SPPM_NEWTON_FAST, abi.SPPM_NEWTON_FAST_BERNSTEIN])
def test_newton_fast_modes_within_1e6(dh, oracle, mode):
    """FAST modes (FMA, reciprocal by rcp + 2 Newton-Raphson steps — here a 20-bit stand-in for
    rcp.approx.ftz.f64 — and Cramer's rule) against the reference arithmetic: same hit/miss,
    (t, u, v) within the north-star 1e-6."""
    sc = make_scene("reference", textures=None)
    P = _patches(sc)
    A = dh.scene(sc).tables()["A"]
    prm = _params(newton_restarts=50)
    rng = np.random.RandomState(9)
    n = 3200
    patch = (np.arange(n) % P.shape[0]).astype(np.int32)
    rays = rays_at_patches(patch, 10, sc)
    seeds = rng.uniform(size=(n, 50, 2))
    ref = oracle.bezier_patch_batch(sc, prm, rays, patch, seeds)
    got = np.empty_like(ref)
    for p in range(P.shape[0]):
        sel = patch == p
        got[sel] = dh.newton(mode, A[p] if mode == abi.SPPM_NEWTON_FAST else P[p], rays[sel], seeds[sel])
    hit = ref[:, 0] < 1e19
    assert hit.sum() > 2000
    np.testing.assert_array_equal(got[:, 0] < 1e19, hit)
    rel = np.abs(got[hit] - ref[hit]) / np.maximum(np.abs(ref[hit]), 1e-12)
    assert rel.max() <= 1e-6  # north-star tolerance on (t, u, v)


# ---- the tables sppm_scene_upload builds (scene_prep.h) ----------------------------------------
@pytest.mark.parametrize("name", ["reference", "stress"])
def test_scene_tables(dh, oracle, name):
    sc = _scene(name)
    P = _patches(sc).reshape(-1, 16, 3)
    hs = dh.scene(sc)
    T = hs.tables()
    assert hs.n_patches == P.shape[0]
    # per-patch and per-object AABBs: the control points' min / max (Bezier.h:165-175, 219-222)
    np.testing.assert_array_equal(T["box"][:, :3], P.min(1))
    np.testing.assert_array_equal(T["box"][:, 3:], P.max(1))
    first = 0
    for i, o in enumerate(sc.objvec):
        if o.kind == 2:
            k = o.P.reshape(-1, 16, 3).shape[0]
            np.testing.assert_array_equal(T["objbox"][i, :3], P[first:first + k].min((0, 1)))
            np.testing.assert_array_equal(T["objbox"][i, 3:], P[first:first + k].max((0, 1)))
            first += k
    # the cull volumes contain every surface point grown by the acceptance radius sqrt(M_EPS) = 0.01
    A13 = np.array([[1, 0, 0], [0, 1, 0], [0, 0, 1], [1, 1, 0], [1, -1, 0], [1, 0, 1], [1, 0, -1],
                    [0, 1, 1], [0, 1, -1], [1, 1, 1], [1, 1, -1], [1, -1, 1], [-1, 1, 1]], dtype=float)
    rng = np.random.RandomState(11)
    for p in range(P.shape[0]):
        uv = rng.uniform(size=(600, 2))
        S, _, _ = dh.bezier_eval(0, P[p].reshape(48), uv)
        e = S @ A13.T
        r = 0.01 * np.linalg.norm(A13, axis=1)
        assert (e - r >= T["dop"][p, :13]).all() and (e + r <= T["dop"][p, 13:]).all()
        b = np.minimum((uv[:, 0] * 4).astype(int), 3) * 4 + np.minimum((uv[:, 1] * 4).astype(int), 3)  # su * 4 + sv
        lo, hi = T["sub"][p][b, :3], T["sub"][p][b, 3:]
        assert (S - 0.01 >= lo).all() and (S + 0.01 <= hi).all()
    # material flags and Roulette weights (Element.h:16-18, Vec3.h:72-79, Obj.h:12-15)
    for i, o in enumerate(sc.objvec):
        fl, w = hs.material(i)
        m = o.material
        for k, ch in enumerate((m.diff, m.refl, m.refr)):
            assert fl[k] == (not any(abs(c) < 1e-4 for c in ch))
            assert w[k] == (ch[0] + ch[1] + ch[2]) / 3
        assert w[3] == w[0] + w[1] + w[2]


def test_quirk_q8_any_zero_channel_is_not_diffuse(dh):
    """IsZero(Vec3) is true when ANY channel is < 1e-4 (Vec3.h:72-79): (.5, 0, .5) is not IsDiff()
    though its Roulette weight is 1/3 (Obj.h:12-15) — prep_scene keeps both facts."""
    hs = dh.scene(quirks_scene())
    fl, w = hs.material(5)
    assert not fl[0] and fl[1] and not fl[2]
    assert w[0] == (0.5 + 0.0 + 0.5) / 3


def test_prep_scene_rejects_bad_scenes(dh):
    sc = make_scene("reference", textures=None)
    d = sc.flatten()
    d.objects[8].first_patch = 5        # 5 + 32 > 32 patches
    with pytest.raises(ValueError, match="patch range"):
        dh.scene(d)
    d = sc.flatten()
    d.objects[0].tex_id = 3
    with pytest.raises(ValueError, match="tex_id"):
        dh.scene(d)
    d = sc.flatten()
    d.objects[1].kind = 7
    with pytest.raises(ValueError, match="kind"):
        dh.scene(d)


def test_newton_culls_never_drop_a_solve_that_records(dh, oracle):
    """Soundness of the two Newton culls (13-DOP line test, sub-patch boxes): whenever the
    reference's solve records a root on a patch, neither cull may fire for that (ray, patch)."""
    sc = make_scene("reference", textures=None)
    hs = dh.scene(sc)
    prm = _params(newton_restarts=50)
    n_p = hs.n_patches
    n = 40000
    patch = (np.arange(n) % n_p).astype(np.int32)
    rays = rays_at_patches(patch, 12, sc)
    k = n // 2
    rays[:k] = rays_at_teapot(k, 13, sc)       # rays that graze or miss their patch
    ref = oracle.bezier_patch_batch(sc, prm, rays, patch, None)
    dop, sub = hs.newton_cull(rays, patch)
    hit = ref[:, 0] < 1e19
    assert hit.sum() > 15000
    assert not (dop & hit).any()
    assert not (sub & hit).any()
    miss = ~hit
    assert (dop | sub)[miss].mean() > 0.5      # and they do prove most of the empty solves empty


def test_analytic_rejects_do_not_change_the_winner(dh, oracle):
    """test_plane / test_sphere (Obj.h:65-85, 111-139) with the sign and FP32-estimate shortcuts:
    the same nearest object and t as the reference's divide-everything loop."""
    sc = make_scene("cornell", textures=None)
    hs = dh.scene(sc)
    rays = np.concatenate([rays_in_room(60000, 14), rays_in_room(200, 15)])
    rays[-200:, 3:] = np.eye(3)[np.arange(200) % 3] * np.where(np.arange(200) % 2, 1.0, -1.0)[:, None]  # axis-parallel
    ref = oracle.intersect_batch(sc, _params(), rays)
    obj, t, inside = hs.analytic_nearest(rays)
    np.testing.assert_array_equal(obj, ref["obj"])
    np.testing.assert_array_equal(t, ref["t"])
    np.testing.assert_array_equal(inside, ref["inside"])


# ---- Scene::NearCollide: intersect_kernel (near_collide + shade_hit), one-warp blocks on fibers ----
@pytest.mark.parametrize("name", ["cornell", "custom", "reference", "stress", "quirks"])
def test_intersect_kernel_golden_of_the_reference(dh, name):
    """near_collide + shade_hit (scene.cuh) == Scene::NearCollide as computed by the reference's own
    PlaneObj / SphereObj / Bezier3Obj classes (tests/golden/ref_*near_collide.npz)."""
    fn = "ref_quirks_near_collide.npz" if name == "quirks" else "ref_near_collide.npz"
    g = np.load(os.path.join(GOLDEN, fn))
    hs = dh.scene(_scene(name))
    rays = g[name + "_rays"]
    got = hs.near_collide(_params(), rays, np.arange(rays.shape[0], dtype=np.uint64), pass_id=1, node=2)
    np.testing.assert_array_equal(got["obj"], g[name + "_obj"])
    np.testing.assert_array_equal(got["inside"], g[name + "_inside"])
    np.testing.assert_array_equal(got["t"], g[name + "_t"])
    np.testing.assert_array_equal(got["pos"], g[name + "_pos"])
    np.testing.assert_allclose(got["n"], g[name + "_n"], rtol=0, atol=1e-15)
    np.testing.assert_allclose(got["color"], g[name + "_color"], rtol=1e-12, atol=1e-15)


@pytest.mark.parametrize("name,n_room,n_tea", [("reference", 20000, 12000), ("stress", 4000, 2500), ("quirks", 4000, 4000)])
def test_intersect_kernel_fuzz_against_oracle(dh, oracle, name, n_room, n_tea):
    """A fuzz some 6x the GPU test's size: queue flushes, ragged last warp, culls on and off."""
    sc = _scene(name)
    hs = dh.scene(sc)
    prm = _params()
    rays = np.concatenate([rays_in_room(n_room, 16), rays_at_teapot(n_tea, 17, sc)])[:-5]  # ragged warp
    ids = np.arange(rays.shape[0], dtype=np.uint64) * 7 + 3
    ref = oracle.intersect_batch(sc, prm, rays, ids, pass_id=2, node=5)
    got = hs.near_collide(prm, rays, ids, pass_id=2, node=5)
    for k in ("obj", "patch", "inside", "t", "pos", "u", "v"):
        np.testing.assert_array_equal(got[k], ref[k], err_msg=k)
    np.testing.assert_allclose(got["n"], ref["n"], rtol=0, atol=1e-15)
    np.testing.assert_allclose(got["color"], ref["color"], rtol=1e-12, atol=1e-15)
    assert (ref["patch"] >= 0).sum() > 500 and hs.culled > 1000
    sub = slice(0, None, 9)
    nocull = hs.near_collide(prm, rays[sub], ids[sub], pass_id=2, node=5, cull=0)
    for k in ("obj", "patch", "t", "u", "v"):
        np.testing.assert_array_equal(nocull[k], ref[k][sub], err_msg=k)


def test_intersect_kernel_fast_newton_within_1e6(dh, oracle):
    """The bench configuration (SPPM_NEWTON_FAST) picks the same object and patch wherever the
    runner-up is not within 1e-6, and lands within 1e-6 on t and the hit position."""
    sc = make_scene("reference", textures="synthetic")
    hs = dh.scene(sc)
    rays = np.concatenate([rays_in_room(6000, 18), rays_at_teapot(6000, 19, sc)])
    ids = np.arange(rays.shape[0], dtype=np.uint64)
    ref = oracle.intersect_batch(sc, _params(), rays, ids, pass_id=3, node=1)
    got = hs.near_collide(_params(newton_mode=abi.SPPM_NEWTON_FAST), rays, ids, pass_id=3, node=1)
    same = (got["obj"] == ref["obj"]) & (got["patch"] == ref["patch"])
    assert same.mean() > 0.9995
    hit = same & (ref["obj"] >= 0)
    assert (np.abs(got["t"][hit] - ref["t"][hit]) <= 1e-6 * np.abs(ref["t"][hit])).all()
    assert np.abs(got["pos"][hit] - ref["pos"][hit]).max() <= 1e-6 * 250


# ---- the KERNELS of trace_kernels.cuh, launched as one-warp blocks -------------------------------
# Same libm on both sides here, so what the GPU tests can only hold to 99.8 % (acos / sin / cos
# differ by ulps between CUDA and libm) is held bit for bit: every hit point, every deposit.
def test_patch_batch_kernel(dh, oracle):
    """patch_batch_kernel (the M2 microbenchmark's kernel): the reference's golden (t, u, v), and the
    oracle with injected and with counter-RNG seeds; grid-stride loop with fewer warps than work."""
    sc = make_scene("reference", textures=None)
    hs = dh.scene(sc)
    g = np.load(os.path.join(GOLDEN, "ref_newton.npz"))
    prm = _params()
    np.testing.assert_array_equal(hs.patch_batch_kernel(prm, g["rays"], g["patch"], g["seeds"], blocks=3), g["tuv"])
    n = 1500
    rng = np.random.RandomState(21)
    patch = rng.randint(0, 32, size=n).astype(np.int32)
    rays = rays_at_patches(patch, 22, sc)
    rays[:300] = rays_at_teapot(300, 23, sc)
    ref = oracle.bezier_patch_batch(sc, prm, rays, patch, None)
    np.testing.assert_array_equal(hs.patch_batch_kernel(prm, rays, patch, None, blocks=7), ref)
    np.testing.assert_array_equal(hs.patch_batch_kernel(prm, rays, patch, None, cull=1, blocks=16), ref)
    assert hs.culled > 100 and (ref[:, 0] < 1e19).sum() > 700


@pytest.mark.parametrize("name,w,h", [("reference", 320, 240), ("custom", 160, 120), ("quirks", 160, 120), ("stress", 128, 96)])
def test_eye_pass_kernel_against_oracle(dh, oracle, name, w, h):
    """eye_pass_kernel == GetHitPoint / ViewTrace (Raytracer.h:102-116, 296-357) of the oracle for every
    pixel — the reference scene at BASELINE config 1's full 320x240 — incl. branching materials
    (deferred refraction subtrees) and the row-sharded launch of the multi-GPU eye pass."""
    from raytrace3_b200.scene import reference_camera
    sc = _scene(name)
    hs = dh.scene(sc)
    prm = _params()
    cam = reference_camera(w, h)
    ref = oracle.eye_pass(sc, prm, cam, pass_id=3)
    got, cnt = hs.eye_pass_kernel(prm, cam, pass_id=3)
    order = np.lexsort((ref["node"], ref["pix"][:, 0] * w + ref["pix"][:, 1]))
    assert got.shape[0] == order.shape[0]
    np.testing.assert_array_equal(got["row"], ref["pix"][order, 0])
    np.testing.assert_array_equal(got["col"], ref["pix"][order, 1])
    np.testing.assert_array_equal(got["okey"] & np.uint64((1 << 20) - 1), ref["node"][order])
    for k in ("pos", "n", "wgt", "r2"):
        np.testing.assert_array_equal(got[k], ref[k][order], err_msg=k)
    assert cnt["isect_queries"] >= w * h
    if name == "custom":
        assert got.shape[0] > w * h * 1.02      # branching
    cut = h // 3
    a, _ = hs.eye_pass_kernel(prm, cam, pass_id=3, row_begin=0, row_end=cut)
    b, _ = hs.eye_pass_kernel(prm, cam, pass_id=3, row_begin=cut, row_end=h)
    assert np.array_equal(np.concatenate([a, b]), got)


def test_eye_pass_kernel_branching_overflow(dh, oracle):
    """About 8 hit points per pixel (conftest.branching_scene): the kernel reports the true count when its
    buffer is too small (nothing is written past the capacity), and the second pass holds them all."""
    from conftest import branching_scene
    from raytrace3_b200.scene import reference_camera
    sc = branching_scene()
    hs = dh.scene(sc)
    prm = _params()
    w, h = 48, 36
    cam = reference_camera(w, h)
    ref = oracle.eye_pass(sc, prm, cam, pass_id=1)
    assert ref["pos"].shape[0] > 4 * w * h + 1024
    got, _ = hs.eye_pass_kernel(prm, cam, pass_id=1)
    order = np.lexsort((ref["node"], ref["pix"][:, 0] * w + ref["pix"][:, 1]))
    assert got.shape[0] == order.shape[0]
    assert len(np.unique(got["okey"])) == got.shape[0]       # (pixel, node) identifies a hit point
    for k in ("pos", "n", "wgt", "r2"):
        np.testing.assert_array_equal(got[k], ref[k][order], err_msg=k)


@pytest.mark.parametrize("name", ["custom", "quirks", "reference"])
def test_eye_pass_kernel_golden_of_the_reference(dh, name):
    """The hit points the reference's OWN recursion produced (tests/golden/ref_*view_trace.npz)."""
    from raytrace3_b200.scene import reference_camera
    fn = "ref_quirks_view_trace.npz" if name == "quirks" else "ref_view_trace.npz"
    g = np.load(os.path.join(GOLDEN, fn))
    w, h = (int(v) for v in g["film"])
    got, _ = dh.scene(_scene(name)).eye_pass_kernel(_params(), reference_camera(w, h), pass_id=0)
    gp = got["row"].astype(np.int64) * w + got["col"]
    off = 0
    for (i, j), cnt in zip(g[name + "_pixels"], g[name + "_counts"]):
        sel = np.nonzero(gp == int(i) * w + int(j))[0]
        assert sel.size == cnt, (i, j)
        want = {f: g[f"{name}_{f}"][off:off + cnt] for f in ("pos", "n", "wgt", "r2")}
        off += cnt
        a = sel[np.lexsort(got["pos"][sel].T[::-1])]      # recursion order vs (pixel, node) order
        b = np.lexsort(want["pos"].T[::-1])
        for f in want:
            np.testing.assert_array_equal(got[f][a], want[f][b], err_msg=f)
    assert off == g[name + "_pos"].shape[0]


@pytest.mark.parametrize("name,count", [("reference", 100000), ("custom", 20000), ("quirks", 20000), ("stress", 20000)])
def test_photon_trace_kernel_against_oracle(dh, oracle, name, count):
    """photon_trace_kernel == Light::emit + PhotonTrace (Light.h:8-13, Raytracer.h:117-209) of the
    oracle: the same deposit list (x, N, incoming flux per (photon, bounce)) and the same counters,
    bit for bit — the reference scene with BASELINE config 1's 100 k photons."""
    sc = _scene(name)
    hs = dh.scene(sc)
    prm = _params()
    ref = oracle.photon_pass(sc, prm, 100, count, 9, keep_deposits=True)
    dep, cnt = hs.photon_trace_kernel(prm, 100, count, 9)
    rd, rc = ref["dep"], ref["counters"]
    assert cnt["photons"] == rc["photons"] == count * len(sc.plvec)
    assert cnt["isect_queries"] == rc["queries"] and cnt["deposits"] == rc["deposits"] == cnt["deposits_kept"]
    assert cnt["patch_solves"] == rc["solves"]
    order = np.lexsort((rd["bounce"], rd["photon"]))
    assert dep["x"].shape[0] == order.shape[0] > 2 * count
    for k in ("photon", "bounce", "x", "n", "flux"):
        np.testing.assert_array_equal(dep[k], rd[k][order], err_msg=k)


def test_photon_trace_kernel_is_launch_invariant(dh):
    """Persistent warps pick photons from a global cursor: the deposits do not depend on the grid."""
    hs = dh.scene(_scene("reference"))
    prm = _params()
    one, _ = hs.photon_trace_kernel(prm, 7, 3001, 4, blocks=1)
    many, _ = hs.photon_trace_kernel(prm, 7, 3001, 4, blocks=40)
    for k in one:
        np.testing.assert_array_equal(one[k], many[k], err_msg=k)
    # photon-range invariance: [7, 3008) = [7, 1000) + [1000, 3008), ids are global
    a, _ = hs.photon_trace_kernel(prm, 7, 993, 4)
    b, _ = hs.photon_trace_kernel(prm, 1000, 2008, 4)
    for k in one:
        np.testing.assert_array_equal(np.concatenate([a[k], b[k]]), one[k], err_msg=k)


@pytest.mark.parametrize("name", ["cornell", "custom", "quirks", "reference"])
def test_photon_trace_kernel_golden_of_the_reference(dh, oracle, name):
    """Level 2 against the reference itself: the kernel's deposits of 300 photons, gathered over the
    golden hit points by the exhaustive search, give the newN / tao the reference's own PhotonTrace +
    radiusSearch accumulated (tests/golden/ref_*photon.npz) — bit for bit, per photon and in total."""
    fn = "ref_quirks_photon.npz" if name == "quirks" else "ref_photon.npz"
    g = np.load(os.path.join(GOLDEN, fn))
    hp = {k: g[f"{name}_hp_{k}"] for k in ("pos", "n", "wgt", "r2", "pix")}
    prm = _params()
    count = g[f"{name}_tot_newN"].shape[0]
    dep, _ = dh.scene(_scene(name)).photon_trace_kernel(prm, 0, count, 5)
    n = hp["pos"].shape[0]
    acc_n, acc_t = np.zeros(n, np.int64), np.zeros((n, 3))
    search_r2 = float(hp["r2"].max())
    for j in range(count):
        sel = (dep["photon"] & np.uint64((1 << 40) - 1)) == j
        d = {k: dep[k][sel] for k in ("x", "n", "flux")}
        newN, tao, _, _ = oracle.gather_brute(hp, d, search_r2)
        assert int(newN.sum()) == g[f"{name}_tot_newN"][j], j
        assert np.array_equal(tao.sum(axis=0), g[f"{name}_tot_tao"][j]), j
        acc_n += newN
        acc_t += tao
    assert np.array_equal(acc_n, g[f"{name}_acc_newN"]) and np.array_equal(acc_t, g[f"{name}_acc_tao"])


# ---- the bounce-synchronous pass (wavefront_kernels.cuh): the bench's default tracer -------------
@pytest.mark.parametrize("name,count", [("reference", 100000), ("custom", 20000), ("quirks", 20000), ("stress", 20000),
                                        ("cornell", 20000)])
def test_wavefront_kernels_against_oracle(dh, oracle, name, count):
    """photon_step_kernel + newton_items_kernel per bounce, driven by the product's own wf_layout /
    wf_step_args / wf_newton_args, == Light::emit + PhotonTrace of the oracle: every deposit and every
    counter, bit for bit (reference scene: BASELINE config 1's 100 k photons).  Also checks the
    four-region queue: each bounce's survivors = the next bounce's input, all photons accounted for."""
    sc = _scene(name)
    hs = dh.scene(sc)
    prm = _params()
    ref = oracle.photon_pass(sc, prm, 100, count, 9, keep_deposits=True)
    dep, cnt, qc = hs.photon_trace_wavefront(prm, 100, count, 9)
    rd, rc = ref["dep"], ref["counters"]
    total = count * len(sc.plvec)
    assert cnt["photons"] == rc["photons"] == total
    assert cnt["deposits"] == rc["deposits"] == cnt["deposits_kept"]
    assert cnt["patch_solves"] == rc["solves"]
    # the oracle counts a query per NearCollide call; the wavefront drops a photon that hits nothing when
    # it finds that out (same query), and never issues the query past the depth limit (Raytracer.h:129-132)
    assert rc["queries"] - qc[prm.max_dep + 2].sum() <= cnt["isect_queries"] <= rc["queries"]
    order = np.lexsort((rd["bounce"], rd["photon"]))
    assert dep["x"].shape[0] == order.shape[0] > 2 * total
    for k in ("photon", "bounce", "x", "n", "flux"):
        np.testing.assert_array_equal(dep[k], rd[k][order], err_msg=k)
    alive = qc.sum(axis=1).astype(np.int64)
    assert alive[0] == 0 and alive[1] <= total and (np.diff(alive[1:]) <= 0).all()
    if any(o.kind == 2 for o in sc.objvec):
        assert qc[1:, 3].sum() > 0          # photons with Bezier solves pending
    else:
        assert qc[:, 3].sum() == 0


@pytest.mark.parametrize("tail", [-1, 6, 3, 0])
def test_wavefront_equals_persistent_on_the_cpu(dh, tail):
    """Both tracer organisations, FAST Newton + FAST math (the bench configuration): same deposits — with the
    tail hand-over off (-1, the product's setting), late (6), early (3) and with every Bezier solve handed over (0);
    odd values also switch the Newton kernel to dynamic chunks."""
    hs = dh.scene(_scene("stress"))
    prm = _params(newton_mode=abi.SPPM_NEWTON_FAST, math_mode=abi.SPPM_MATH_FAST)
    a, ca = hs.photon_trace_kernel(prm, 0, 30000, 2)
    b, cb, _ = hs.photon_trace_wavefront(prm, 0, 30000, 2, step_blocks=5, newton_blocks=3, tail_bounce=tail)
    for k in a:
        np.testing.assert_array_equal(a[k], b[k], err_msg=k)
    assert ca["deposits"] == cb["deposits"] and ca["patch_solves"] == cb["patch_solves"]
    assert ca["isect_queries"] >= cb["isect_queries"] > 0.9 * ca["isect_queries"]
    assert (cb["overflow"] > 0) == (tail >= 0)


def _light_over_teapot_scene():
    """The case the (ray, patch) item list can be too short for (2 items per photon and bounce): the light hangs
    just above the teapot's bounding box (AABBbox::intersect only sees rays that ENTER a box, Q5, so a light inside
    it queues nothing), and every downward ray passes the object box and several patch boxes: 1.5 solves per
    emitted photon instead of the shipped scene's 0.037."""
    from raytrace3_b200.scene import make_scene
    sc = make_scene("reference", textures=None)
    bz = [o for o in sc.objvec if o.kind == 2][0]
    lo, hi = bz.P.reshape(-1, 3).min(0), bz.P.reshape(-1, 3).max(0)
    sc.plvec[0].pos = tuple(0.5 * (lo + hi) + np.array([0.0, 0.6 * (hi[1] - lo[1]), 0.0]))
    return sc


@pytest.mark.parametrize("name,item_cap", [("reference", 16), ("reference", 1), ("stress", 64), ("light_over_teapot", 4096),
                                           ("light_over_teapot", 0)])
def test_wavefront_item_list_spills_to_the_persistent_tracer(dh, oracle, name, item_cap):
    """A bounce whose item list is full hands its photons to photon_trace_kernel's resume mode instead of dropping
    their Bezier solves (Bezier.h:240-282 must run for every boxed ray): deposits and counters equal the oracle's,
    bit for bit, however short the list; item_cap = 0 is the product's sizing."""
    sc = _light_over_teapot_scene() if name == "light_over_teapot" else _scene(name)
    hs = dh.scene(sc)
    prm = _params()
    count = 6000
    ref = oracle.photon_pass(sc, prm, 0, count, 4, keep_deposits=True)
    dep, cnt, qc = hs.photon_trace_wavefront(prm, 0, count, 4, step_blocks=7, newton_blocks=5, item_cap=item_cap)
    rd, rc = ref["dep"], ref["counters"]
    order = np.lexsort((rd["bounce"], rd["photon"]))
    for k in ("photon", "bounce", "x", "n", "flux"):
        np.testing.assert_array_equal(dep[k], rd[k][order], err_msg=k)
    assert cnt["photons"] == rc["photons"] and cnt["deposits"] == rc["deposits"]
    assert cnt["patch_solves"] == rc["solves"]
    if item_cap:
        assert cnt["overflow"] > 0          # photons did spill
    if name == "light_over_teapot":
        assert rc["solves"] > count


# ---- the hash grid and the FP32-filtered gather (gather_kernels.cuh) -----------------------------
# build_index's kernels (cell keys, SoA permute, neighbourhood slots, work units), the deposit keys
# and gather_filtered_kernel run as one-warp blocks; CUB's sorts / scans are std:: algorithms in the
# harness.  The inputs are those of the GPU tests of the same names (tests/test_gpu_parity.py).
def _check_gather(dh, oracle, hp, x, dn, flux, search_r2=2.0, min_acc=100):
    """FILTERED (default) and EXACT gather kernels against the exhaustive search."""
    dep = dict(x=x, n=dn, flux=flux)
    newN, tao, cand, acc = oracle.gather_brute(hp, dep, search_r2)
    assert newN.sum() > min_acc
    # count_pairs = False is the build the bench times (gather_filtered_kernel<COUNT = false>): the cull radius follows
    # the unit's largest R2 and units whose hit points share one normal test it once per deposit (cand_pairs stays 0)
    for exact, count in ((False, True), (False, False), (True, True)):
        gn, gt, info, pairs = dh.gather_filtered(hp, dep, search_r2, exact=exact, count_pairs=count)
        np.testing.assert_array_equal(gn, newN)                 # level 2: newN bit-exact
        assert pairs[1] == acc and (pairs[0] == cand or not count), (exact, count)  # cand_pairs is only kept by the COUNT build
        np.testing.assert_allclose(gt, tao, rtol=1e-11, atol=0)  # far inside the 1e-5 bar: only the order of summation differs
    return info


def test_gather_filtered_plane_units_and_shrunk_radii(dh, oracle):
    """Hit points on planes (one normal per plane, bit for bit: the units of a wall) with radii that PROGRESSIVE updates
    have shrunk (R2 from 0.05 to 2): the normal test of Raytracer.h:154 is made once per deposit and unit, the cull
    radius is the unit's largest R2.  Deposits sit at n.N = 1e-3 (1 +- delta) and at distance^2 = R2 (1 +- delta) of a
    hit point, delta down to 1e-16: newN must still equal the exhaustive FP64 search."""
    rng = np.random.RandomState(21)
    sheets = [((0., 1., 0.), (0, 2)), ((1., 0., 0.), (1, 2)), ((0.6, 0., 0.8), (0, 1))]
    pos, nrm = [], []
    for nvec, (a, b) in sheets:
        k = 2500
        p = np.zeros((k, 3))
        p[:, a], p[:, b] = rng.uniform(0, 12, k), rng.uniform(0, 12, k)
        free = 3 - a - b
        p[:, free] = 3.0 + 1e-3 * rng.uniform(-1, 1, k)
        pos.append(p)
        nrm.append(np.tile(np.array(nvec), (k, 1)))
    pos, nrm = np.concatenate(pos), np.concatenate(nrm)
    n = pos.shape[0]
    r2 = rng.choice([0.05, 0.3, 0.9, 1.4, 2.0], size=n)
    r2[:2500][pos[:2500, 0] < 4] = 0.3           # whole units with one small radius
    hp = dict(pos=pos, n=nrm, wgt=rng.uniform(0.1, 1, size=(n, 3)), r2=r2)
    m = 40000
    who = rng.randint(0, n, m)
    delta = rng.choice([0.0, 1e-16, 1e-15, 1e-13, 1e-10, 1e-7, 1e-4, 1e-2], size=m) * rng.choice([-1, 1], size=m)
    d = rng.normal(size=(m, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    rad = np.sqrt(r2[who] * (1 + delta))
    half = m // 2
    rad[half:] = np.sqrt(r2[who[half:]]) * rng.uniform(0, 1.3, m - half)
    x = pos[who] + d * rad[:, None]
    # normals at the agreement threshold of the hit point's plane
    cosv = np.where(rng.uniform(size=m) < 0.5, 1e-3 * (1 + delta), rng.uniform(-1, 1, m))
    t = np.cross(nrm[who], rng.normal(size=(m, 3)))
    t /= np.linalg.norm(t, axis=1, keepdims=True)
    dn = nrm[who] * cosv[:, None] + t * np.sqrt(np.maximum(0, 1 - cosv ** 2))[:, None]
    info = _check_gather(dh, oracle, hp, x, dn, rng.uniform(0, 100, size=(m, 3)), min_acc=10000)
    assert info["n_units"] > 100


def test_gather_filtered_whole_unit_accept_on_dense_planes(dh, oracle):
    """A dense hit-point map (the 4K film's regime: a work unit of 64 hit points is much smaller than the radius):
    deposits inside the ball of EVERY hit point of a unit are accepted for the whole unit without a pair test.  Deposits
    at distance^2 = T (1 +- delta) of single hit points, T in {searchR2, R2}, delta down to 1e-16, and at the normal
    threshold must still give the exhaustive search's newN bit for bit; the full-accept path must have been taken."""
    rng = np.random.RandomState(33)
    k = 9000
    pos = np.zeros((k, 3))
    pos[:, 0], pos[:, 2] = rng.uniform(0, 5, k), rng.uniform(0, 5, k)
    pos[:, 1] = 1.0
    nrm = np.tile(np.array([0., 1., 0.]), (k, 1))
    side = np.zeros((3000, 3))                       # a second plane meeting the first at a right angle
    side[:, 1], side[:, 2] = rng.uniform(1, 4, 3000), rng.uniform(0, 5, 3000)
    pos = np.concatenate([pos, side])
    nrm = np.concatenate([nrm, np.tile(np.array([1., 0., 0.]), (3000, 1))])
    n = pos.shape[0]
    r2 = np.where(rng.uniform(size=n) < 0.5, 2.0, rng.choice([0.6, 1.1, 1.7], size=n))
    r2[(pos[:, 0] > 2.5) & (pos[:, 1] == 1.0)] = 2.0   # half of the floor: units with R2 == searchR2 throughout
    hp = dict(pos=pos, n=nrm, wgt=rng.uniform(0.1, 1, size=(n, 3)), r2=r2)
    m = 30000
    who = rng.randint(0, n, m)
    delta = rng.choice([0.0, 1e-16, 1e-15, 1e-13, 1e-10, 1e-7, 1e-4], size=m) * rng.choice([-1, 1], size=m)
    d = rng.normal(size=(m, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    T = np.where(rng.uniform(size=m) < 0.5, 2.0, r2[who])
    rad = np.sqrt(T * (1 + delta))
    rad[m // 2:] = rng.uniform(0, 1.6, m - m // 2)
    x = pos[who] + d * rad[:, None]
    x[m // 2:, 1] = np.where(rng.uniform(size=m - m // 2) < 0.7, 1.0 + 1e-3 * rng.normal(size=m - m // 2), x[m // 2:, 1])  # most lie on the floor
    cosv = np.where(rng.uniform(size=m) < 0.2, 1e-3 * (1 + delta), rng.uniform(0.01, 1, m))
    t = np.cross(nrm[who], rng.normal(size=(m, 3)))
    t /= np.linalg.norm(t, axis=1, keepdims=True)
    dn = nrm[who] * cosv[:, None] + t * np.sqrt(np.maximum(0, 1 - cosv ** 2))[:, None]
    dh.gather_work(reset=True)
    _check_gather(dh, oracle, hp, x, dn, rng.uniform(0, 100, size=(m, 3)), min_acc=100000)
    assert dh.gather_work()["full_accepts"] > 10000


def test_gather_filtered_kernel_bit_exact_newN(dh, oracle):
    """Raytracer.h:139-159 incl. deposits exactly at distance^2 == searchR2 (strict '<', Q3) and == R2 ('<=')."""
    rng = np.random.RandomState(7)
    n, m = 6000, 20000
    hp = dict(pos=np.stack([rng.uniform(0, 30, n), rng.uniform(0, 2, n), rng.uniform(0, 30, n)], axis=1),
              wgt=rng.uniform(0.1, 1, size=(n, 3)), r2=rng.uniform(0.2, 2.0, size=n))
    nrm = rng.normal(size=(n, 3))
    hp["n"] = nrm / np.linalg.norm(nrm, axis=1, keepdims=True)
    x = np.stack([rng.uniform(-2, 32, m), rng.uniform(-1, 3, m), rng.uniform(-2, 32, m)], axis=1)
    dn = rng.normal(size=(m, 3))
    dn /= np.linalg.norm(dn, axis=1, keepdims=True)
    flux = rng.uniform(0, 100, size=(m, 3))
    hp["pos"][0], hp["r2"][0], hp["n"][0] = (10.0, 1.0, 10.0), 2.0, (0, 1, 0)
    x[0] = (11.0, 2.0, 10.0)   # vv == 2.0 exactly -> excluded by the strict radius search
    x[1] = (11.0, 1.0, 10.0)   # vv == 1.0 -> included
    hp["pos"][1], hp["r2"][1], hp["n"][1] = (20.0, 1.0, 20.0), 1.0, (0, 1, 0)
    x[2] = (21.0, 1.0, 20.0)   # vv == R2 exactly -> accepted ('<=')
    dn[0] = dn[1] = dn[2] = (0, 1, 0)
    info = _check_gather(dh, oracle, hp, x, dn, flux, min_acc=1000)
    assert info["n_units"] >= info["n_occ"] // 40 and info["kept"] > m // 2


def test_gather_filtered_kernel_threshold_band_adversarial(dh, oracle):
    """Deposits at dist^2 = T(1+d) for T in {searchR2, R2}, |d| from 1e-16 to 1e-4 (both signs, and 0),
    normals with n.N = 1e-3(1+d), coordinates near 150 where an FP32 ulp is 1.5e-5: the FP32
    classification must fall back to FP64 inside its error bound, never misclassify."""
    rng = np.random.RandomState(11)
    n_hp, m = 3000, 60000
    pos = np.stack([rng.uniform(120, 150, n_hp), rng.uniform(100, 103, n_hp), rng.uniform(130, 150, n_hp)], axis=1)
    nrm = rng.normal(size=(n_hp, 3))
    nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
    r2 = np.where(rng.uniform(size=n_hp) < 0.5, 2.0, rng.uniform(0.3, 2.0, size=n_hp))
    hp = dict(pos=pos, n=nrm, wgt=rng.uniform(0.1, 1, size=(n_hp, 3)), r2=r2)
    who = rng.randint(0, n_hp, m)
    e = rng.normal(size=(m, 3))
    e /= np.linalg.norm(e, axis=1, keepdims=True)
    delta = np.sign(rng.uniform(-1, 1, m)) * 10.0 ** rng.uniform(-16, -4, m)
    delta[rng.uniform(size=m) < 0.05] = 0.0
    T = np.where(rng.uniform(size=m) < 0.5, 2.0, r2[who])
    x = pos[who] + e * np.sqrt(T * (1 + delta))[:, None]
    t = np.cross(nrm[who], rng.normal(size=(m, 3)))
    t /= np.linalg.norm(t, axis=1, keepdims=True)
    d2 = np.sign(rng.uniform(-1, 1, m)) * 10.0 ** rng.uniform(-16, -3, m)
    cosv = np.where(rng.uniform(size=m) < 0.5, 1e-3 * (1 + d2), rng.uniform(0.2, 1.0, m))
    dn = nrm[who] * cosv[:, None] + t * np.sqrt(1 - cosv ** 2)[:, None]
    near = rng.uniform(size=m) < 0.3
    x[near] = pos[who[near]] + e[near] * rng.uniform(0.1, 1.0, near.sum())[:, None]
    _check_gather(dh, oracle, hp, x, dn, rng.uniform(0, 100, size=(m, 3)), min_acc=10000)


@pytest.mark.parametrize("seed", range(4))
def test_gather_filtered_kernel_shapes_fuzz(dh, oracle, seed):
    """Planar sheets, a sparse run along x, one cell with hundreds of hit points (units split every 64),
    isolated points, thousands of deposits on a few cells (units cut into several work items, which
    accumulate with atomics), deposits far outside the hit points' bounds, per-hit-point radii."""
    rng = np.random.RandomState(100 + seed)
    n_sheet = rng.randint(200, 4000)
    parts = [np.stack([rng.uniform(0, 40, n_sheet), np.full(n_sheet, 3.0) + rng.normal(0, 1e-3, n_sheet), rng.uniform(0, 25, n_sheet)], 1)]
    n_run = rng.randint(5, 80)
    parts.append(np.stack([50 + 1.42 * np.arange(n_run) + rng.uniform(0, 0.3, n_run), np.full(n_run, 10.0), np.full(n_run, 5.0)], 1))
    n_blob = rng.randint(70, 600)
    parts.append(np.array([20.0, 20.0, 20.0]) + rng.uniform(-0.6, 0.6, size=(n_blob, 3)))
    parts.append(rng.uniform(-30, 120, size=(rng.randint(1, 30), 3)))
    pos = np.concatenate(parts)
    n = pos.shape[0]
    nrm = rng.normal(size=(n, 3))
    nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
    nrm[:n_sheet] = (0, 1, 0)
    hp = dict(pos=pos, n=nrm, wgt=rng.uniform(0.1, 1, size=(n, 3)),
              r2=np.where(rng.uniform(size=n) < 0.6, 2.0, rng.uniform(0.05, 2.0, size=n)))
    m = rng.randint(2000, 60000)
    x = pos[rng.randint(0, n, m)] + rng.normal(0, rng.choice([0.3, 1.0, 3.0]), size=(m, 3))
    heavy = rng.randint(0, m, m // 3)
    x[heavy] = np.array([20.0, 20.0, 20.0]) + rng.normal(0, 1.0, size=(heavy.size, 3))
    x[: m // 50] = rng.uniform(-5000, 5000, size=(m // 50, 3))
    dn = rng.normal(size=(m, 3))
    dn /= np.linalg.norm(dn, axis=1, keepdims=True)
    dn[rng.uniform(size=m) < 0.5] = (0, 1, 0)
    info = _check_gather(dh, oracle, hp, x, dn, rng.uniform(0, 50, size=(m, 3)))
    assert info["kept"] < m                      # the far deposits are not binned
    if m > 12000:
        assert info["n_items"] > info["n_units"] - info["n_occ"]  # some unit was cut into several work items


def test_bounds_and_canonical_order_kernels(dh, oracle):
    """bounds_kernel (min / max / sum through the order-preserving integer keys, negative coordinates
    included) and hp_okey_kernel: hit points handed over in arrival order give the same map."""
    rng = np.random.RandomState(61)
    n, m = 3000, 9000
    pos = rng.uniform(-40, 60, size=(n, 3))
    lo, hi, sm = dh.bounds(pos)
    np.testing.assert_array_equal(lo, pos.min(0))
    np.testing.assert_array_equal(hi, pos.max(0))
    np.testing.assert_allclose(sm, pos.sum(0), rtol=1e-12)
    hp = dict(pos=rng.uniform(0, 14, size=(n, 3)), wgt=rng.uniform(0.1, 1, size=(n, 3)), r2=rng.uniform(0.5, 2.0, n))
    nrm = rng.normal(size=(n, 3))
    hp["n"] = nrm / np.linalg.norm(nrm, axis=1, keepdims=True)
    dep = dict(x=rng.uniform(-1, 15, size=(m, 3)), n=hp["n"][rng.randint(0, n, m)], flux=rng.uniform(0, 9, size=(m, 3)))
    a = dh.gather_filtered(hp, dep, 2.0)
    b = dh.gather_filtered(hp, dep, 2.0, shuffle_seed=3)
    np.testing.assert_array_equal(a[0], b[0])
    np.testing.assert_allclose(a[1], b[1], rtol=1e-12, atol=0)   # the deposits' arrival ranks within a cell vary from run to run
    assert a[3] == b[3] and a[0].sum() > 1000


def test_gather_filtered_kernel_out_of_range_normals_take_the_exact_path(dh, oracle):
    """A deposit normal with a component above 1.0001 (not a unit vector: the ABI takes any deposit list)
    voids the FP32 error bound; such a slice is evaluated in FP64 wholesale."""
    rng = np.random.RandomState(5)
    n, m = 800, 5000
    hp = dict(pos=rng.uniform(0, 12, size=(n, 3)), wgt=rng.uniform(0.1, 1, size=(n, 3)), r2=np.full(n, 2.0))
    nrm = rng.normal(size=(n, 3))
    hp["n"] = nrm / np.linalg.norm(nrm, axis=1, keepdims=True)
    x = rng.uniform(-1, 13, size=(m, 3))
    dn = rng.normal(size=(m, 3)) * 3.0           # |components| up to ~10
    _check_gather(dh, oracle, hp, x, dn, rng.uniform(0, 10, size=(m, 3)))


@pytest.mark.parametrize("name", ["cornell", "reference"])
def test_photon_kernels_and_gather_kernel_reproduce_the_reference(dh, name):
    """The whole device path on the CPU against the reference's own code: wavefront kernels -> deposits
    -> grid + filtered gather kernel over the golden hit points == the newN the reference's PhotonTrace
    + FLANN radius search accumulated over 300 photons (bit-exact), tao to rounding."""
    g = np.load(os.path.join(GOLDEN, "ref_photon.npz"))
    hp = {k: g[f"{name}_hp_{k}"] for k in ("pos", "n", "wgt", "r2")}
    count = g[f"{name}_tot_newN"].shape[0]
    dep, _, _ = dh.scene(_scene(name)).photon_trace_wavefront(_params(), 0, count, 5)
    newN, tao, _, _ = dh.gather_filtered(hp, dep, float(hp["r2"].max()))
    np.testing.assert_array_equal(newN, g[f"{name}_acc_newN"])
    np.testing.assert_allclose(tao, g[f"{name}_acc_tao"], rtol=1e-12, atol=0)
    assert newN.sum() > 50


# ---- random scenes ------------------------------------------------------------------------------
def _random_scene(seed):
    """Scenes the fixed ones never combine: oblique planes, spheres of any size (the light may sit inside
    one), materials with any mix of diffuse / mirror / glass incl. zero channels and odd indices of
    refraction, textures on any object, mirror / glass / textured teapots under random transforms."""
    from raytrace3_b200.scene import (Bezier3Obj, Material, PlaneObj, Scene, SphereObj, SpotLight, Texture, teapot_patches)
    rng = np.random.RandomState(1000 + seed)

    branching = [rng.uniform() < 0.6]   # at most one object whose eye paths branch (mirror AND glass: 2^depth hit points)

    def material(may_branch=False):
        ch = lambda: tuple(float(v) for v in rng.choice([0.0, 0.0, 0.3, 0.6, 0.999], 3)) if rng.uniform() < 0.3 else \
            (float(rng.uniform(0.1, 0.999)),) * 3
        kind = rng.randint(0, 5)
        if may_branch and branching[0]:
            kind, branching[0] = 5, False
        Z = (0.0, 0.0, 0.0)
        d, s, r = [(ch(), Z, Z), (Z, ch(), Z), (Z, Z, ch()), (ch(), ch(), Z), (ch(), Z, ch()), (ch(), ch(), ch())][kind]
        return Material(d, s, r, float(rng.choice([1.5, 1.33, 2.4, 0.8])))

    def colour():
        return tuple(float(v) for v in rng.uniform(0.2, 0.999, 3))

    def texture():
        return Texture(synthetic_texture(int(rng.randint(3, 90)), int(rng.randint(3, 90)), int(rng.randint(1 << 20))))

    s = Scene()
    s.objvec = []
    walls = [((0, 40, 80), (1, 0, 0)), ((100, 40, 80), (1, 0, 0)), ((50, 40, 0), (0, 0, 1)), ((50, 0, 80), (0, -1, 0)),
             ((50, 82, 80), (0, 1, 0)), ((50, 40, 230), (0, 0, -1))]
    for p, nrm in walls:
        if rng.uniform() < 0.85:
            nrm = np.array(nrm, dtype=float)
            if rng.uniform() < 0.4:
                nrm = nrm + rng.normal(0, 0.25, 3)        # oblique: Obj.h:89-96 picks the texture axes by the last non-zero component
            s.objvec.append(PlaneObj(p, tuple(nrm), colour(), material()))
    for _ in range(rng.randint(1, 6)):
        s.objvec.append(SphereObj(float(rng.uniform(2, 25)), tuple(rng.uniform((10, 5, 20), (90, 60, 200))), colour(), material(True)))
    if rng.uniform() < 0.3:                               # a sphere around the light
        s.objvec.append(SphereObj(float(rng.uniform(5, 40)), (50, 60, 85), colour(), material()))
    tp = teapot_patches()
    for _ in range(rng.randint(0, 3)):
        q, _r = np.linalg.qr(rng.normal(size=(3, 3)))
        s.objvec.append(Bezier3Obj(tuple(rng.uniform((15, 0, 40), (85, 40, 180))), float(rng.uniform(1.5, 5)), tp, colour(),
                                   material(), q))
    for o in s.objvec:
        if rng.uniform() < 0.4:
            o.setTexture(texture())
    rng.shuffle(s.objvec)                                 # Scene::NearCollide: first object wins ties
    s.plvec = [SpotLight((50, 60, 85), (5000., 4000., 3000.))]
    if rng.uniform() < 0.4:
        s.plvec.append(SpotLight(tuple(rng.uniform((10, 10, 20), (90, 70, 200))), (100., 2000., 900.)))
    return s


@pytest.mark.parametrize("seed", range(24))
def test_random_scenes_eye_and_photon_kernels(dh, oracle, seed):
    from raytrace3_b200.scene import reference_camera
    sc = _random_scene(seed)
    hs = dh.scene(sc)
    prm = _params()
    cam = reference_camera(80, 60)
    ref = oracle.eye_pass(sc, prm, cam, pass_id=seed)
    got, _ = hs.eye_pass_kernel(prm, cam, pass_id=seed)
    order = np.lexsort((ref["node"], ref["pix"][:, 0] * 80 + ref["pix"][:, 1]))
    assert got.shape[0] == order.shape[0]
    np.testing.assert_array_equal(got["okey"] & np.uint64((1 << 20) - 1), ref["node"][order])
    for k in ("pos", "n", "wgt", "r2"):
        np.testing.assert_array_equal(got[k], ref[k][order], err_msg=k)
    count = 3000
    rp = oracle.photon_pass(sc, prm, 11, count, seed, keep_deposits=True)
    rd = rp["dep"]
    order = np.lexsort((rd["bounce"], rd["photon"]))
    for name, (dep, cnt) in (("wavefront", hs.photon_trace_wavefront(prm, 11, count, seed)[:2]),
                             ("persistent", hs.photon_trace_kernel(prm, 11, count, seed))):
        assert cnt["deposits"] == rp["counters"]["deposits"], name
        if name == "persistent":
            assert cnt["patch_solves"] == rp["counters"]["solves"] and cnt["isect_queries"] == rp["counters"]["queries"]
        else:   # the wavefront never issues the query past the depth limit, whose result nobody reads (Raytracer.h:129-132)
            assert cnt["patch_solves"] <= rp["counters"]["solves"]
        for k in ("photon", "bounce", "x", "n", "flux"):
            np.testing.assert_array_equal(dep[k], rd[k][order], err_msg=f"{name} {k}")


@pytest.mark.parametrize("mode", [abi.SPPM_UPDATE_REFERENCE, abi.SPPM_UPDATE_PROGRESSIVE])
def test_update_and_resolve_kernels(dh, oracle, mode):
    """update_kernel == HitPoint::update (Raytracer.h:69-79: the shipped no-op while N == 0 (Q1), the
    progressive radius / flux shrink with the int truncation of N) and resolve_kernel == the film
    expression of Raytracer.h:281-293 bit for bit (a pixel's hit points are summed in Hvec order), three
    rounds of fresh accumulators."""
    rng = np.random.RandomState(31)
    n, w, h = 5000, 40, 30
    N, newN = np.zeros(n, np.int32), np.zeros(n, np.int32)
    N[: n // 2] = rng.randint(0, 50, n // 2)             # REFERENCE mode only acts where N > 0
    r2, tao = rng.uniform(0.2, 2.0, n), np.zeros((n, 3))
    pix = np.stack([rng.randint(0, h, n), rng.randint(0, w, n)], axis=1).astype(np.int32)
    oN, onewN, or2, otao = N.copy(), newN.copy(), r2.copy(), tao.copy()
    for _ in range(3):
        dN = rng.randint(0, 4, n).astype(np.int32)
        dtao = rng.uniform(0, 10, size=(n, 3)) * (dN > 0)[:, None]
        onewN += dN                                      # the round's accumulators, then HitPoint::update
        otao += dtao
        oracle.update(mode, 0.7, onewN, oN, or2, otao)
        tao, film = dh.update_resolve(mode, 0.7, N, newN, dN, r2, tao, dtao, pix, 3, 1000, w, h)   # zeroes dN / dtao
        assert not dN.any()
    np.testing.assert_array_equal(N, oN)
    np.testing.assert_array_equal(newN, onewN)
    np.testing.assert_array_equal(r2, or2)
    np.testing.assert_array_equal(tao, otao)
    want = oracle.resolve(pix, or2, otao, 3, 1000, w, h)
    np.testing.assert_array_equal(film, want)   # several hit points per pixel: summed in canonical order, as the oracle
    assert (N != 0).any() and film.sum() > 0


def test_resolve_kernel_with_a_pixels_hit_points_apart_in_canonical_order(dh, oracle):
    """An imported map numbers its hit points in the caller's order (sppm_import_hitpoints: okey = index), so a
    pixel's hit points need not be neighbours in canonical order: resolve_kernel then adds one partial sum per run
    of neighbours — the same film as the oracle's up to the order of those adds."""
    rng = np.random.RandomState(32)
    n, w, h = 3000, 20, 15                      # 10 hit points per pixel on average, scattered over the index range
    N, newN, dN = np.zeros(n, np.int32), rng.randint(1, 9, n).astype(np.int32), np.zeros(n, np.int32)
    r2, tao = rng.uniform(0.2, 2.0, n), rng.uniform(0, 10, size=(n, 3))
    pix = np.stack([rng.randint(0, h, n), rng.randint(0, w, n)], axis=1).astype(np.int32)
    _, film = dh.update_resolve(abi.SPPM_UPDATE_REFERENCE, 0.7, N, newN, dN, r2, tao.copy(), np.zeros((n, 3)), pix, 2, 500, w, h,
                                rounds=0, canonical="index")
    want = oracle.resolve(pix, r2, tao, 2, 500, w, h)
    np.testing.assert_allclose(film, want, rtol=1e-13, atol=0)
    assert film.sum() > 0


def test_shared_statistics_kernels(dh, oracle):
    """shared_{init,load,accum,update,resolve}_kernel: three eye passes of the per-pixel statistics
    (N' = N + alpha M, R2' = R2 (N + alpha M)/(N + M), tau' = (tau + Phi) R2'/R2; Hachisuka & Jensen 2009)
    against the restated fold (oracle.shared_fold); radii only shrink, hit points measure with their
    pixel's current radius."""
    from raytrace3_b200.scene import PI
    rng = np.random.RandomState(41)
    w, h = 24, 18
    r2, n, tau = np.zeros((h, w)), np.zeros((h, w)), np.zeros((h, w, 3))
    o_r2, o_n, o_tau = np.full((h, w), 2.0), np.zeros((h, w)), np.zeros((h, w, 3))
    photons = 0
    for k in range(3):
        n_hp = 700 + 50 * k                                # hit points move with the jittered camera
        pix = np.stack([rng.randint(0, h, n_hp), rng.randint(0, w, n_hp)], axis=1).astype(np.int32)
        newN = rng.randint(0, 5, n_hp).astype(np.int32)
        tao = rng.uniform(0, 10, size=(n_hp, 3)) * (newN > 0)[:, None]
        photons += 1000
        want_r2 = o_r2[pix[:, 0], pix[:, 1]].copy()        # before the fold
        hp_r2, film = dh.shared_pass(r2, n, tau, pix, newN, tao, 0.7, float(photons), init_r2=2.0 if k == 0 else 0.0)
        oracle.shared_fold(o_r2, o_n, o_tau, pix, newN, tao, 0.7)
        np.testing.assert_array_equal(hp_r2, want_r2)
        np.testing.assert_allclose(r2, o_r2, rtol=1e-14, atol=0)
        np.testing.assert_allclose(n, o_n, rtol=1e-14, atol=0)
        np.testing.assert_allclose(tau, o_tau, rtol=1e-13, atol=0)
        np.testing.assert_allclose(film, o_tau / (PI * o_r2 * photons)[..., None], rtol=1e-12, atol=0)   # the reference's PI (Vec3.h:8)
    assert (r2 <= 2.0).all() and (r2 < 2.0).any()


def test_tonemap_and_texel_expansion_kernels(dh, oracle):
    """tonemap_kernel == toInt + the row flip of Raytracer.h:24-26, 460-470 (same libm: bit for bit, incl. the
    values that land on a .5 boundary region and the clamp-free overflow of huge radiances);
    tex_expand_kernel == B | G << 8 | R << 16 per texel."""
    rng = np.random.RandomState(51)
    film = rng.uniform(0, 30, size=(37, 53, 3)) * (rng.uniform(size=(37, 53, 1)) < 0.9)
    film[0, :5] = [[0, 0, 0], [1e-12, 1e-6, 1e-3], [4.0, 40.0, 400.0], [1e5, 1e9, 1e12], [0.5, 2.0, 8.0]]
    for div in (4.0, 8.0, 400.0):
        np.testing.assert_array_equal(dh.tonemap(film, div), oracle.tonemap(film, div))
    tex = synthetic_texture(19, 45, 3)
    want = tex[..., 0].astype(np.uint32) | (tex[..., 1].astype(np.uint32) << 8) | (tex[..., 2].astype(np.uint32) << 16)
    levels = dh.tex_expand(tex)
    np.testing.assert_array_equal(levels[0], want)      # level 0 (8 x 4 tiles on the device) == the uploaded image


def test_texture_mip_chain_and_level_sampling(dh, oracle):
    """SURVEY 8f-4 (texture tiling / mip): level l+1 = the rounded 2 x 2 box mean of level l (an odd edge repeats),
    down to 1 x 1; color_uv at lod 0 is Texture::colorUV of the reference bit for bit (tiled storage is
    invisible), at lod l it is the same wrap-around bilinear rule on level l."""
    rng = np.random.RandomState(52)
    tex = rng.randint(0, 256, size=(37, 70, 3)).astype(np.uint8)
    levels = dh.tex_expand(tex)
    cur = tex.astype(np.int64)
    assert len(levels) == 7 and levels[-1].shape == (1, 1)
    for l, got in enumerate(levels):
        packed = cur[..., 0] | (cur[..., 1] << 8) | (cur[..., 2] << 16)
        np.testing.assert_array_equal(got, packed.astype(np.uint32), err_msg=f"level {l}")
        r, c = cur.shape[:2]
        nr, nc = max(1, r >> 1), max(1, c >> 1)
        ri = np.minimum(np.arange(nr)[:, None] * 2 + np.array([0, 1]), r - 1)      # [nr, 2]
        ci = np.minimum(np.arange(nc)[:, None] * 2 + np.array([0, 1]), c - 1)
        blk = cur[ri[:, None, :, None], ci[None, :, None, :]]                        # [nr, nc, 2, 2, 3]
        cur = (blk.sum(axis=(2, 3)) + 2) >> 2
    uv = rng.uniform(-2, 3, size=(4000, 2))
    want0 = np.array([oracle.color_uv(tex, u, v) for u, v in uv])
    np.testing.assert_array_equal(dh.color_uv(tex, uv, lod=0), want0)
    lvl2 = np.stack([(levels[2] >> (8 * k)) & 255 for k in range(3)], axis=-1).astype(np.uint8)
    want2 = np.array([oracle.color_uv(lvl2, u, v) for u, v in uv[:1500]])
    np.testing.assert_array_equal(dh.color_uv(tex, uv[:1500], lod=2), want2)


def test_one_warp_blocks_give_the_same_results(dh, oracle):
    """The tests above launch the kernels with the product's block sizes (128-thread trace / step blocks,
    256-thread gather blocks, 320-thread scatter); one-warp blocks are a second schedule of the same
    work (other warp-to-work assignment, other arrival order of the deposits)."""
    sc = _scene("quirks")
    hs = dh.scene(sc)
    prm = _params()
    from raytrace3_b200.scene import reference_camera
    cam = reference_camera(96, 72)
    eye, _ = hs.eye_pass_kernel(prm, cam, pass_id=2)
    dep, cnt, _ = hs.photon_trace_wavefront(prm, 5, 8000, 2)
    hp = dict(pos=eye["pos"], n=eye["n"], wgt=eye["wgt"], r2=eye["r2"])
    newN, tao, _, pairs = dh.gather_filtered(hp, dep, 2.0)
    dh.one_warp_blocks(True)
    try:
        eye1, _ = hs.eye_pass_kernel(prm, cam, pass_id=2)
        dep1, cnt1, _ = hs.photon_trace_wavefront(prm, 5, 8000, 2)
        per1, _ = hs.photon_trace_kernel(prm, 5, 8000, 2)
        newN1, tao1, _, pairs1 = dh.gather_filtered(hp, dep, 2.0)
    finally:
        dh.one_warp_blocks(False)
    assert np.array_equal(eye, eye1) and cnt == cnt1 and pairs == pairs1
    for k in dep:
        np.testing.assert_array_equal(dep[k], dep1[k], err_msg=k)
        np.testing.assert_array_equal(dep[k], per1[k], err_msg=k)
    np.testing.assert_array_equal(newN, newN1)
    np.testing.assert_allclose(tao, tao1, rtol=1e-12, atol=0)
    assert newN.sum() > 1000


def test_device_kernels_are_clean_under_asan_ubsan():
    """A memcheck of the kernel bodies without a GPU: the same emulation library built with
    -fsanitize=address,undefined (SPPM_DEVICE_HOST_SANITIZE=1) runs a cross-section of this file — every
    kernel of trace / wavefront / gather_kernels.cuh on several scenes — without a single report (global
    and shared-memory bounds, alignment, shifts, signed overflow).  The whole file was run this way for
    profiles/r1_K_device_host_asan.txt and, at the end of round 2, profiles/r2_S_device_host_asan.txt."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if os.environ.get("SPPM_DEVICE_HOST_SANITIZE"):
        pytest.skip("already inside the sanitizer run")
    asan = subprocess.run(["g++", "-print-file-name=libasan.so"], capture_output=True, text=True).stdout.strip()
    if not os.path.isabs(asan) or not os.path.exists(asan):
        pytest.skip("libasan.so not installed")
    env = dict(os.environ, LD_PRELOAD=asan, SPPM_DEVICE_HOST_SANITIZE="1", UBSAN_OPTIONS="print_stacktrace=1",
               ASAN_OPTIONS="detect_leaks=0:halt_on_error=1:detect_stack_use_after_return=0")
    pick = ("intersect_kernel_golden or patch_batch or eye_pass_kernel_branching or "
            "(wavefront_kernels_against_oracle and stress) or (photon_trace_kernel_golden and quirks) or launch_invariant or "
            "(shapes_fuzz and 0) or out_of_range_normals or (reproduce_the_reference and cornell) or "
            "(random_scenes and 7) or update_and_resolve or (item_list_spills and light_over_teapot) or whole_unit_accept or "
            "plane_units or texture_mip_chain")
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.abspath(__file__), "-q", "-s", "-x", "-p", "no:cacheprovider",
                        "-k", pick], capture_output=True, text=True, env=env, cwd=root, timeout=1500)
    out = r.stdout + r.stderr
    assert r.returncode == 0, out[-3000:]
    assert "runtime error" not in out and "ERROR: AddressSanitizer" not in out, out[-3000:]
    assert " passed" in out
