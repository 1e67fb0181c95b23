"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle on the same
seeded inputs.  Bars (BASELINE.json north_star): intersection t/u/v <= 1e-6 relative (EXACT
mode: bit-identical), gather newN bit-exact and tao <= 1e-5 relative, image rRMSE <= 1%."""
import numpy as np
import pytest

from conftest import rays_at_patches, rays_at_teapot, rays_in_room
from raytrace3_b200 import abi
from raytrace3_b200.scene import make_scene, reference_camera

pytestmark = pytest.mark.gpu


def _params(**kw):
    p = abi.default_params()
    for k, v in kw.items():
        setattr(p, k, v)
    return p


def test_patch_newton_injected_seeds_exact(ctx, oracle):
    """Bezier3::newton (Bezier.h:112-159) with injected (u0,v0): EXACT mode is bit-identical."""
    sc = make_scene("reference", textures=None)
    prm = _params(newton_mode=abi.SPPM_NEWTON_EXACT)
    ctx.set_params(prm)
    ctx.upload_scene(sc)
    n = 700
    rng = np.random.RandomState(5)
    patch = rng.randint(0, 32, size=n).astype(np.int32)
    rays = rays_at_patches(patch, 11, sc)
    rays[:100] = rays_at_teapot(100, 11, sc)  # and some that mostly miss their patch
    seeds = rng.uniform(size=(n, prm.newton_restarts, 2))
    ref = oracle.bezier_patch_batch(sc, prm, rays, patch, seeds)
    got = ctx.bezier_patch_batch(rays, patch, seeds)
    assert (ref[:, 0] < 1e19).sum() > 300 and (ref[:, 0] >= 1e19).sum() > 50  # solutions and misses
    np.testing.assert_array_equal(got, ref)


@pytest.mark.parametrize("mode", [abi.SPPM_NEWTON_FAST, abi.SPPM_NEWTON_FAST_BERNSTEIN])
def test_patch_newton_fast_within_1e6(ctx, oracle, mode):
    """FAST (monomial Horner + FMA, the bench default) and FAST_BERNSTEIN against the oracle."""
    sc = make_scene("reference", textures=None)
    prm = _params(newton_mode=mode)
    ctx.set_params(prm)
    ctx.upload_scene(sc)
    n = 3000
    patch = (np.arange(n) % 32).astype(np.int32)
    rays = rays_at_patches(patch, 12, sc)
    ref = oracle.bezier_patch_batch(sc, prm, rays, patch, None)  # counter-RNG seeds
    got = ctx.bezier_patch_batch(rays, patch, None)
    hit_r, hit_g = ref[:, 0] < 1e19, got[:, 0] < 1e19
    assert hit_r.sum() > 1500
    assert (hit_r == hit_g).all()
    rel = np.abs(got[hit_r] - ref[hit_r]) / np.maximum(np.abs(ref[hit_r]), 1e-12)
    assert rel.max() <= 1e-6  # north-star tolerance on (t, u, v)


@pytest.mark.parametrize("scene_name,textured", [("cornell", False), ("reference", False), ("reference", True), ("stress", False)])
def test_near_collide_matches_oracle(ctx, oracle, scene_name, textured):
    """Scene::NearCollide (Scene.h:165-182): same object, same hit/miss, same t/Pos/N/colour."""
    sc = make_scene(scene_name, textures="synthetic" if textured else None)
    prm = _params()
    ctx.set_params(prm)
    ctx.upload_scene(sc)
    rays = np.concatenate([rays_in_room(4000, 3), rays_at_teapot(1500, 4, sc) if scene_name != "cornell" else rays_in_room(500, 9)])
    ids = np.arange(rays.shape[0], dtype=np.uint64) * 7 + 3
    ref = oracle.intersect_batch(sc, prm, rays, ids, pass_id=2, node=5)
    got = ctx.intersect_batch(rays, ids, pass_id=2, node=5)
    np.testing.assert_array_equal(got["obj"], ref["obj"])
    np.testing.assert_array_equal(got["patch"], ref["patch"])
    np.testing.assert_array_equal(got["inside"], ref["inside"])
    np.testing.assert_array_equal(got["t"], ref["t"])  # +,-,*,/,sqrt only: bit-exact
    np.testing.assert_array_equal(got["pos"], ref["pos"])
    np.testing.assert_array_equal(got["u"], ref["u"])
    np.testing.assert_array_equal(got["v"], ref["v"])
    np.testing.assert_allclose(got["n"], ref["n"], rtol=0, atol=1e-15)
    np.testing.assert_allclose(got["color"], ref["color"], rtol=1e-12, atol=1e-15)  # sphere texture uses acos/sin
    if scene_name != "cornell":
        assert (ref["patch"] >= 0).sum() > 100


def test_eye_pass_hitpoints(ctx, oracle):
    """GetHitPoint/ViewTrace (Raytracer.h:102-116, 296-357) incl. mirror + glass chains."""
    sc = make_scene("reference", textures="synthetic")
    prm = _params()
    ctx.set_params(prm)
    ctx.upload_scene(sc)
    cam = reference_camera(96, 72)
    ref = oracle.eye_pass(sc, prm, cam, pass_id=3)
    n = ctx.eye_pass(cam, pass_id=3)
    assert n == ref["pos"].shape[0]
    got = ctx.export_hitpoints()
    order = np.lexsort((ref["node"], ref["pix"][:, 0] * cam.w + ref["pix"][:, 1]))  # canonical: pixel, node
    for k in ("pix", "node"):
        np.testing.assert_array_equal(got[k], ref[k][order])
    np.testing.assert_array_equal(got["pos"], ref["pos"][order])
    np.testing.assert_allclose(got["n"], ref["n"][order], rtol=0, atol=1e-15)
    np.testing.assert_allclose(got["wgt"], ref["wgt"][order], rtol=1e-12, atol=0)
    np.testing.assert_array_equal(got["r2"], ref["r2"][order])


def _random_hitpoints(rng, n, w=64, h=48):
    pos = np.stack([rng.uniform(0, 30, n), rng.uniform(0, 2, n), rng.uniform(0, 30, n)], axis=1)
    nrm = rng.normal(size=(n, 3))
    nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
    wgt = rng.uniform(0.1, 1, size=(n, 3))
    r2 = rng.uniform(0.2, 2.0, size=n)
    pix = np.stack([rng.randint(0, h, n), rng.randint(0, w, n)], axis=1).astype(np.int32)
    return dict(pos=pos, n=nrm, wgt=wgt, r2=r2, pix=pix), w, h


GATHER_MODES = [abi.SPPM_GATHER_EXACT, abi.SPPM_GATHER_FILTERED]


@pytest.mark.parametrize("gmode", GATHER_MODES)
def test_gather_bit_exact_newN(ctx, oracle, gmode):
    """Raytracer.h:139-159: newN bit-exact, tao <= 1e-5 relative vs the exhaustive oracle,
    incl. deposits exactly at distance^2 == searchR2 (strict '<', Q3) and == R2 ('<=')."""
    rng = np.random.RandomState(7)
    hp, w, h = _random_hitpoints(rng, 6000)
    m = 20000
    x = np.stack([rng.uniform(-2, 32, m), rng.uniform(-1, 3, m), rng.uniform(-2, 32, m)], axis=1)
    dn = rng.normal(size=(m, 3))
    dn /= np.linalg.norm(dn, axis=1, keepdims=True)
    flux = rng.uniform(0, 100, size=(m, 3))
    # boundary cases: hit point 0 at the origin of a 3-4-5 style exact offset
    hp["pos"][0] = (10.0, 1.0, 10.0)
    hp["r2"][0] = 2.0
    hp["n"][0] = (0, 1, 0)
    x[0] = (11.0, 2.0, 10.0)   # vv == 2.0 exactly -> excluded by the strict radius search
    x[1] = (11.0, 1.0, 10.0)   # vv == 1.0 -> included
    hp["pos"][1] = (20.0, 1.0, 20.0)
    hp["r2"][1] = 1.0          # vv == R2 exactly -> accepted ('<=')
    hp["n"][1] = (0, 1, 0)
    x[2] = (21.0, 1.0, 20.0)
    dn[0] = dn[1] = dn[2] = (0, 1, 0)
    prm = _params(gather_mode=gmode)
    ctx.set_params(prm)
    ctx.import_hitpoints(hp["pos"], hp["n"], hp["wgt"], hp["r2"], hp["pix"], w, h)
    ctx.build_index()
    ctx.reset_counters()
    ctx.gather_deposits(x, dn, flux)
    ctx.update()
    got = ctx.export_stats(6000)
    newN, tao, cand, acc = oracle.gather_brute(hp, dict(x=x, n=dn, flux=flux), prm.init_r2)
    assert newN.sum() > 1000
    np.testing.assert_array_equal(got["newN"], newN)
    np.testing.assert_allclose(got["tao"], tao, rtol=1e-5, atol=0)
    c = ctx.counters()
    assert c["cand_pairs"] == cand and c["acc_pairs"] == acc
    # a second batch accumulates on top (Q1: newN is never reset in REFERENCE mode)
    ctx.gather_deposits(x[:5000], dn[:5000], flux[:5000])
    ctx.update()
    got2 = ctx.export_stats(6000)
    newN2, tao2, _, _ = oracle.gather_brute(hp, dict(x=x[:5000], n=dn[:5000], flux=flux[:5000]), prm.init_r2, newN.copy(), tao.copy())
    np.testing.assert_array_equal(got2["newN"], newN2)
    np.testing.assert_allclose(got2["tao"], tao2, rtol=1e-5, atol=0)


@pytest.mark.parametrize("gmode", GATHER_MODES)
def test_gather_edge_cases(ctx, oracle, gmode):
    prm = _params(gather_mode=gmode)
    ctx.set_params(prm)
    rng = np.random.RandomState(1)
    # empty hit-point map, empty deposit list, single far-away hit point, clustered cell
    ctx.import_hitpoints(np.zeros((0, 3)), np.zeros((0, 3)), np.zeros((0, 3)), np.zeros(0), np.zeros((0, 2), np.int32), 4, 4)
    ctx.build_index()
    ctx.gather_deposits(rng.uniform(size=(10, 3)), rng.uniform(size=(10, 3)), rng.uniform(size=(10, 3)))
    ctx.update()
    hp, w, h = _random_hitpoints(rng, 300)
    hp["pos"][:200] = hp["pos"][0] + rng.uniform(-0.2, 0.2, size=(200, 3))  # > 32 hit points in one cell
    hp["pos"][299] = (1e7, -1e7, 3e6)                                        # clamped into a boundary cell
    ctx.import_hitpoints(hp["pos"], hp["n"], hp["wgt"], hp["r2"], hp["pix"], w, h)
    ctx.build_index()
    ctx.gather_deposits(np.zeros((0, 3)), np.zeros((0, 3)), np.zeros((0, 3)))
    x = np.concatenate([hp["pos"][:250] + rng.uniform(-0.5, 0.5, size=(250, 3)), [[1e7 + 0.5, -1e7, 3e6]]])
    dn = np.repeat(hp["n"][:1], x.shape[0], axis=0)
    dn[-1] = hp["n"][299]
    flux = rng.uniform(1, 2, size=x.shape)
    ctx.gather_deposits(x, dn, flux)
    ctx.update()
    got = ctx.export_stats(300)
    newN, tao, _, _ = oracle.gather_brute(hp, dict(x=x, n=dn, flux=flux), prm.init_r2)
    assert newN[299] == 1 and newN[:200].max() > 32
    np.testing.assert_array_equal(got["newN"], newN)
    np.testing.assert_allclose(got["tao"], tao, rtol=1e-5, atol=0)


@pytest.mark.parametrize("gmode", GATHER_MODES)
def test_gather_threshold_band_adversarial(ctx, oracle, gmode):
    """The FILTERED gather classifies pairs in FP32 and must fall back to FP64 inside its error
    bound: deposits placed at dist^2 = T(1+d) for T in {searchR2, R2} and |d| from 1e-16 to
    1e-4 (both signs, and d = 0), normals with n.N = 1e-3(1+d), coordinates near 150 where an
    FP32 ulp is 1.5e-5.  newN and the pair counters must equal the FP64 oracle exactly."""
    rng = np.random.RandomState(11)
    n_hp, m = 3000, 60000
    pos = np.stack([rng.uniform(120, 150, n_hp), rng.uniform(100, 103, n_hp), rng.uniform(130, 150, n_hp)], axis=1)
    nrm = rng.normal(size=(n_hp, 3))
    nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
    r2 = np.where(rng.uniform(size=n_hp) < 0.5, 2.0, rng.uniform(0.3, 2.0, size=n_hp))
    hp = dict(pos=pos, n=nrm, wgt=rng.uniform(0.1, 1, size=(n_hp, 3)), r2=r2,
              pix=np.stack([rng.randint(0, 48, n_hp), rng.randint(0, 64, n_hp)], axis=1).astype(np.int32))
    who = rng.randint(0, n_hp, m)
    e = rng.normal(size=(m, 3))
    e /= np.linalg.norm(e, axis=1, keepdims=True)
    delta = np.sign(rng.uniform(-1, 1, m)) * 10.0 ** rng.uniform(-16, -4, m)
    delta[rng.uniform(size=m) < 0.05] = 0.0
    T = np.where(rng.uniform(size=m) < 0.5, 2.0, r2[who])
    x = pos[who] + e * np.sqrt(T * (1 + delta))[:, None]
    # normals: half agree clearly, half sit at the 1e-3 threshold
    t = np.cross(nrm[who], rng.normal(size=(m, 3)))
    t /= np.linalg.norm(t, axis=1, keepdims=True)
    d2 = np.sign(rng.uniform(-1, 1, m)) * 10.0 ** rng.uniform(-16, -3, m)
    cosv = np.where(rng.uniform(size=m) < 0.5, 1e-3 * (1 + d2), rng.uniform(0.2, 1.0, m))
    dn = nrm[who] * cosv[:, None] + t * np.sqrt(1 - cosv ** 2)[:, None]
    near_dist = rng.uniform(size=m) < 0.3      # some threshold-normal deposits well inside the ball
    x[near_dist] = pos[who[near_dist]] + e[near_dist] * rng.uniform(0.1, 1.0, near_dist.sum())[:, None]
    flux = rng.uniform(0, 100, size=(m, 3))
    prm = _params(gather_mode=gmode)
    ctx.set_params(prm)
    ctx.import_hitpoints(hp["pos"], hp["n"], hp["wgt"], hp["r2"], hp["pix"], 64, 48)
    ctx.build_index()
    ctx.reset_counters()
    ctx.gather_deposits(x, dn, flux)
    ctx.update()
    got = ctx.export_stats(n_hp)
    newN, tao, cand, acc = oracle.gather_brute(hp, dict(x=x, n=dn, flux=flux), prm.init_r2)
    assert newN.sum() > 10000
    np.testing.assert_array_equal(got["newN"], newN)
    np.testing.assert_allclose(got["tao"], tao, rtol=1e-5, atol=0)
    c = ctx.counters()
    assert c["cand_pairs"] == cand and c["acc_pairs"] == acc


@pytest.mark.parametrize("seed", range(6))
def test_gather_filtered_fuzz_against_exact(ctx, seed):
    """Random hit-point / deposit clouds of very different shapes (isolated hit points, long sparse
    runs along x, one cell with hundreds of hit points, planar sheets, thousands of deposits on a few
    cells so that units are split into several work items, deposits far outside the hit points'
    bounds, per-hit-point radii): the FILTERED kernel must reproduce the EXACT kernel's newN and pair
    counters bit for bit and tao to rounding."""
    rng = np.random.RandomState(100 + seed)
    parts = []
    n_sheet = rng.randint(200, 4000)
    parts.append(np.stack([rng.uniform(0, 40, n_sheet), np.full(n_sheet, 3.0) + rng.normal(0, 1e-3, n_sheet), rng.uniform(0, 25, n_sheet)], 1))
    n_run = rng.randint(5, 80)                                   # a sparse run along x: one hit point per cell
    parts.append(np.stack([50 + 1.42 * np.arange(n_run) + rng.uniform(0, 0.3, n_run), np.full(n_run, 10.0), np.full(n_run, 5.0)], 1))
    n_blob = rng.randint(70, 600)                                # one dense cell
    parts.append(np.array([20.0, 20.0, 20.0]) + rng.uniform(-0.6, 0.6, size=(n_blob, 3)))
    parts.append(rng.uniform(-30, 120, size=(rng.randint(1, 30), 3)))  # isolated points
    pos = np.concatenate(parts)
    n = pos.shape[0]
    nrm = rng.normal(size=(n, 3))
    nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
    nrm[:n_sheet] = (0, 1, 0)
    r2 = np.where(rng.uniform(size=n) < 0.6, 2.0, rng.uniform(0.05, 2.0, size=n))
    pix = np.stack([rng.randint(0, 48, n), rng.randint(0, 64, n)], axis=1).astype(np.int32)
    wgt = rng.uniform(0.1, 1, size=(n, 3))
    m = rng.randint(2000, 60000)
    x = pos[rng.randint(0, n, m)] + rng.normal(0, rng.choice([0.3, 1.0, 3.0]), size=(m, 3))
    heavy = rng.randint(0, m, m // 3)                            # a third of them on the dense cell and the sheet corner
    x[heavy] = np.array([20.0, 20.0, 20.0]) + rng.normal(0, 1.0, size=(heavy.size, 3))
    x[: m // 50] = rng.uniform(-5000, 5000, size=(m // 50, 3))   # far outside the hit points' bounds
    dn = rng.normal(size=(m, 3))
    dn /= np.linalg.norm(dn, axis=1, keepdims=True)
    flat = rng.uniform(size=m) < 0.5
    dn[flat] = (0, 1, 0)
    flux = rng.uniform(0, 50, size=(m, 3))
    out = []
    for gmode in GATHER_MODES:
        ctx.set_params(_params(gather_mode=gmode))
        ctx.import_hitpoints(pos, nrm, wgt, r2, pix, 64, 48)
        ctx.build_index()
        ctx.reset_counters()
        ctx.gather_deposits(x, dn, flux)
        ctx.update()
        c = ctx.counters()
        out.append((ctx.export_stats(n), (c["cand_pairs"], c["acc_pairs"])))
    (se, ce), (sf, cf) = out
    assert se["newN"].sum() > 100
    np.testing.assert_array_equal(sf["newN"], se["newN"])
    assert cf == ce
    np.testing.assert_allclose(sf["tao"], se["tao"], rtol=1e-11, atol=0)


@pytest.mark.parametrize("gmode", GATHER_MODES)
@pytest.mark.parametrize("mode", [abi.SPPM_UPDATE_REFERENCE, abi.SPPM_UPDATE_PROGRESSIVE])
def test_update_modes(ctx, oracle, mode, gmode):
    """HitPoint::update (Raytracer.h:69-79): verbatim no-op (Q1) and the progressive variant."""
    rng = np.random.RandomState(3)
    hp, w, h = _random_hitpoints(rng, 2000)
    hp["r2"][:] = 2.0
    prm = _params(update_mode=mode, gather_mode=gmode)
    ctx.set_params(prm)
    ctx.import_hitpoints(hp["pos"], hp["n"], hp["wgt"], hp["r2"], hp["pix"], w, h)
    ctx.build_index()
    newN, N = np.zeros(2000, np.int32), np.zeros(2000, np.int32)
    r2, tao = hp["r2"].copy(), np.zeros((2000, 3))
    for rnd in range(3):
        m = 4000
        x = np.stack([rng.uniform(0, 30, m), rng.uniform(0, 2, m), rng.uniform(0, 30, m)], axis=1)
        dn = hp["n"][rng.randint(0, 2000, m)]
        flux = rng.uniform(0, 10, size=(m, 3))
        ctx.gather_deposits(x, dn, flux)
        ctx.update()
        cur = dict(hp, r2=r2)
        oracle.gather_brute(cur, dict(x=x, n=dn, flux=flux), prm.init_r2, newN, tao)
        oracle.update(mode, prm.alpha, newN, N, r2, tao)
    got = ctx.export_stats(2000)
    np.testing.assert_array_equal(got["newN"], newN)
    np.testing.assert_array_equal(got["N"], N)
    np.testing.assert_allclose(got["r2"], r2, rtol=1e-14)
    np.testing.assert_allclose(got["tao"], tao, rtol=1e-5, atol=0)
    if mode == abi.SPPM_UPDATE_PROGRESSIVE:
        assert (r2 < 2.0).any() and N.max() > 0
    else:
        assert (r2 == 2.0).all() and N.max() == 0


@pytest.mark.parametrize("trace_mode", [abi.SPPM_TRACE_WAVEFRONT, abi.SPPM_TRACE_PERSISTENT])
def test_photon_pass_deposits_and_stats(ctx, oracle, trace_mode):
    """Light::emit + PhotonTrace (Light.h:8-13, Raytracer.h:117-209): same photons -> same
    deposit list (x, N, flux per (photon, bounce)); gathered newN equal, tao <= 1e-5.
    Both tracer organisations (bounce-synchronous wavefront, persistent megakernel)."""
    sc = make_scene("reference", textures="synthetic")
    prm = _params(trace_mode=trace_mode)
    ctx.set_params(prm)
    ctx.upload_scene(sc)
    cam = reference_camera(80, 60)
    ctx.eye_pass(cam, 0)
    ctx.build_index()
    hp = ctx.export_hitpoints()
    count = 6000
    ctx.keep_deposits(True)
    try:
        ctx.reset_counters()
        ctx.photon_pass(100, count, 9)
        ctx.update()
        dep = ctx.export_deposits()
    finally:
        ctx.keep_deposits(False)
    ref = oracle.photon_pass(sc, prm, 100, count, 9, hp=hp, keep_deposits=True)
    c = ctx.counters()
    rc = ref["counters"]
    # transcendental functions (acos/sin/cos) differ by ulps between libm and CUDA, which can
    # flip a discrete decision on a rare path: allow a 0.1% mismatch in the path structure
    assert abs(c["isect_queries"] - rc["queries"]) <= 0.001 * rc["queries"]
    assert abs(c["deposits"] - rc["deposits"]) <= 0.001 * rc["deposits"]
    assert abs(c["patch_solves"] - rc["solves"]) <= 0.002 * rc["solves"] + 2
    assert c["photons"] == rc["photons"] == count
    rd = ref["dep"]
    okey = np.lexsort((rd["bounce"], rd["photon"]))
    key_r = {(int(p), int(b)): i for i, (p, b) in enumerate(zip(rd["photon"], rd["bounce"]))}
    matched = 0
    maxerr = 0.0
    for j in range(dep["x"].shape[0]):
        i = key_r.get((int(dep["photon"][j]), int(dep["bounce"][j])))
        if i is None:
            continue
        err = (np.abs(dep["x"][j] - rd["x"][i]) / np.maximum(1.0, np.abs(rd["x"][i]))).max()
        if err < 1e-6:
            matched += 1
            maxerr = max(maxerr, err)
    assert matched >= 0.998 * rd["x"].shape[0], (matched, rd["x"].shape[0], len(okey))
    assert maxerr < 1e-10  # relative: positions agree to ulps (sin/cos/acos differ by ulps)
    got = ctx.export_stats(hp["pos"].shape[0])
    # statistics: identical deposits up to ulps -> newN equal except where a deposit sits within
    # ulps of a radius boundary; require near-total agreement and close totals
    assert (got["newN"] == ref["newN"]).mean() > 0.995
    assert abs(int(got["newN"].sum()) - int(ref["newN"].sum())) <= 0.002 * ref["newN"].sum()
    np.testing.assert_allclose(got["tao"].sum(0), ref["tao"].sum(0), rtol=2e-3)


def test_gather_of_exported_deposits_is_bit_exact(ctx, oracle):
    """Parity level (2) on a traced pass: feed the GPU's own deposit list to the exhaustive
    oracle gather -> newN bit-exact, tao <= 1e-5."""
    sc = make_scene("cornell", textures=None)
    prm = _params()
    ctx.set_params(prm)
    ctx.upload_scene(sc)
    cam = reference_camera(128, 96)
    ctx.eye_pass(cam, 0)
    ctx.build_index()
    hp = ctx.export_hitpoints()
    ctx.keep_deposits(True)
    try:
        ctx.photon_pass(0, 20000, 1)
        ctx.update()
        dep = ctx.export_deposits()
    finally:
        ctx.keep_deposits(False)
    got = ctx.export_stats(hp["pos"].shape[0])
    newN, tao, _, _ = oracle.gather_brute(hp, dep, prm.init_r2)
    assert newN.sum() > 10000
    np.testing.assert_array_equal(got["newN"], newN)
    np.testing.assert_allclose(got["tao"], tao, rtol=1e-5, atol=0)


def test_image_rrmse(ctx, oracle):
    """Parity level (3): film after matched photon counts, relative RMSE <= 1%."""
    from raytrace3_b200 import RayTracer
    sc = make_scene("reference", textures="synthetic")
    prm = _params()
    ctx.set_params(prm)
    w, h, cnt, onetime = 64, 48, 2, 30000
    cam = reference_camera(w, h)
    rt = RayTracer(ctx)
    rt.camera, rt.scene, rt.cnt, rt.onetime = cam, sc, cnt, onetime
    rt.render()
    ref = oracle.render_replicas(sc, prm, cam.pos, cam.look, w, h, cnt, onetime, threads=1, jitter=0.0)
    film_ref, film = ref["film"], cam.pic
    assert film_ref.max() > 0
    rrmse = np.sqrt(((film - film_ref) ** 2).mean()) / np.sqrt((film_ref ** 2).mean())
    assert rrmse <= 0.01, rrmse
    img = ctx.tonemap(film, 1.0)
    img_ref = oracle.tonemap(film_ref, 1.0)
    assert np.abs(img.astype(int) - img_ref.astype(int)).max() <= 2


@pytest.mark.parametrize("world", [1, 2])
def test_fused_rounds_equal_round_by_round(ctx, world):
    """RayTracer::PhotonMap traces the rounds of one eye pass as one photon pass while the update is
    the reference's no-op (Q1): same photons (ids i*onetime + k under the eye pass's id), so newN is
    bit-identical to the round-by-round loop and tao / the film agree to summation order; the ranks
    of a sharded run cover the same ids."""
    from raytrace3_b200 import RayTracer
    sc = make_scene("reference", textures=None)
    prm = _params()
    ctx.set_params(prm)
    cam = reference_camera(64, 48)
    cnt, onetime = 7, 1501
    res = []
    for fuse in (False, True):
        ctx.reset_counters()
        rt = RayTracer(ctx)
        rt.camera, rt.scene, rt.cnt, rt.onetime, rt.fuse_rounds, rt.pass_base = cam.copy(), sc, cnt, onetime, fuse, 3
        n = rt.GetHitPoint()
        ctx.build_index()
        l0 = ctx.counters()["launches"]             # cumulative over the context's life
        for r in range(world):                      # the ranks of a sharded run, one after the other on this GPU
            rt.rank, rt.world = r, world
            rt.PhotonMap(cnt)                       # the last resolve sees every rank's photons
        c = ctx.counters()
        res.append((ctx.export_stats(n), rt.camera.pic.copy(), c, c["launches"] - l0))
    (a, fa, ca, la), (b, fb, cb, lb) = res
    assert ca["photons"] == cb["photons"] == cnt * onetime * len(sc.plvec)
    assert a["newN"].sum() > 0
    np.testing.assert_array_equal(a["newN"], b["newN"])
    np.testing.assert_allclose(a["tao"], b["tao"], rtol=1e-12, atol=0)
    np.testing.assert_allclose(fa, fb, rtol=1e-12, atol=0)
    assert lb < la                                  # one launch sequence per rank instead of cnt


def test_photon_range_invariance(ctx):
    """§8(e): splitting a round's photons over R ranks gives the same accumulators (newN
    bit-exact, tao to rounding) because the RNG is keyed by the global photon id."""
    sc = make_scene("cornell", textures=None)
    prm = _params()
    ctx.set_params(prm)
    ctx.upload_scene(sc)
    cam = reference_camera(96, 72)
    n = ctx.eye_pass(cam, 0)
    from raytrace3_b200 import photon_range
    res = []
    for world in (1, 3):
        ctx.build_index()
        for r in range(world):
            b, c = photon_range(20001, r, world)
            ctx.photon_pass(b, c, 4)
        ctx.update()
        res.append(ctx.export_stats(n))
    np.testing.assert_array_equal(res[0]["newN"], res[1]["newN"])
    np.testing.assert_allclose(res[0]["tao"], res[1]["tao"], rtol=1e-12)


def test_wavefront_equals_persistent(ctx):
    """The two tracer organisations trace the same photons: identical deposit lists (bitwise:
    same arithmetic, same RNG addresses) and identical gathered statistics."""
    sc = make_scene("stress", textures=None)
    cam = reference_camera(64, 48)
    out = []
    for mode in (abi.SPPM_TRACE_WAVEFRONT, abi.SPPM_TRACE_PERSISTENT):
        ctx.set_params(_params(trace_mode=mode, newton_mode=abi.SPPM_NEWTON_FAST))
        ctx.upload_scene(sc)
        n = ctx.eye_pass(cam, 0)
        ctx.build_index()
        ctx.keep_deposits(True)
        try:
            ctx.reset_counters()
            ctx.photon_pass(7, 30000, 2)
            ctx.update()
            out.append((ctx.export_deposits(), ctx.export_stats(n), ctx.counters()))
        finally:
            ctx.keep_deposits(False)
    (d0, s0, c0), (d1, s1, c1) = out
    for k in ("photon", "bounce", "x", "n", "flux"):
        np.testing.assert_array_equal(d0[k], d1[k])
    np.testing.assert_array_equal(s0["newN"], s1["newN"])
    np.testing.assert_allclose(s0["tao"], s1["tao"], rtol=1e-12)
    for k in ("photons", "isect_queries", "patch_solves", "deposits", "deposits_kept", "cand_pairs", "acc_pairs"):
        assert c0[k] == c1[k], k
    assert c0["patch_solves"] > 1000


def test_fast_math_directions_match_exact(ctx):
    """SPPM_MATH_FAST takes sin/cos of theta = acos(sqrt(U1)) as sqrt(1-U1), sqrt(U1) (Vec3.h:90-98
    without the acos -> sincos round trip): the same photons deposit at the same places to
    rounding (paths are chaotic, so a few may split at an edge after many bounces)."""
    sc = make_scene("reference", textures=None)
    cam = reference_camera(64, 48)
    out = []
    for mm in (abi.SPPM_MATH_EXACT, abi.SPPM_MATH_FAST):
        ctx.set_params(_params(math_mode=mm))
        ctx.upload_scene(sc)
        ctx.eye_pass(cam, 0)
        ctx.build_index()
        ctx.keep_deposits(True)
        try:
            ctx.photon_pass(0, 20000, 3)
            out.append(ctx.export_deposits())
        finally:
            ctx.keep_deposits(False)
    d0, d1 = out
    k0 = {(int(p), int(b)): i for i, (p, b) in enumerate(zip(d0["photon"], d0["bounce"]))}
    idx0, idx1 = [], []
    for j, (p, b) in enumerate(zip(d1["photon"], d1["bounce"])):
        i = k0.get((int(p), int(b)))
        if i is not None:
            idx0.append(i)
            idx1.append(j)
    assert len(idx0) >= 0.995 * max(len(d0["photon"]), len(d1["photon"]))
    err = np.abs(d0["x"][idx0] - d1["x"][idx1]).max(axis=1)
    assert np.mean(err < 1e-8) > 0.995
    # after one sampled direction: rounding of a ~100-unit path; acos() near 1 is ill-conditioned
    # (d theta = eps / theta), so the EXACT chain itself carries ~1e-9 at small theta / grazing hits
    first = d0["bounce"][idx0] == 1
    assert np.abs(d0["x"][idx0][first] - d1["x"][idx1][first]).max() < 1e-7
    assert np.median(np.abs(d0["x"][idx0][first] - d1["x"][idx1][first]).max(axis=1)) < 1e-12


def test_shared_statistics_across_eye_passes(ctx, oracle):
    """per-pixel R2 / N / tau persist while the camera jitters.  (a) After one eye
    pass the estimate equals the reference's film (N = 0: R2' = alpha R2, tau' = alpha Phi, so
    tau'/(PI R2' Ne) = Phi/(PI R2 Ne) = Raytracer.h:288-293).  (b) Over three jittered passes
    the device statistics follow the restated update fed with the exported per-pass hit-point
    counters.  (c) Radii only shrink and every hit point measures with its pixel's radius."""
    from raytrace3_b200.rng import unit_rand_vec, P_JITTER
    sc = make_scene("reference", textures=None)
    cam0 = reference_camera(64, 48)
    prm = _params(newton_mode=abi.SPPM_NEWTON_FAST)
    ctx.set_params(prm)
    ctx.upload_scene(sc)
    w, h, onetime, rounds = 64, 48, 20000, 2
    ctx.shared_begin(w, h)
    r2 = np.full((h, w), prm.init_r2)
    n = np.zeros((h, w))
    tau = np.zeros((h, w, 3))
    emitted = 0
    for p in range(3):
        cam = cam0.copy()
        cam.MoveTo(cam0.pos + 0.00015 * unit_rand_vec(prm.seed, 0, p, P_JITTER))
        cam.lookAt(cam0.look, 1.)
        nh = ctx.eye_pass(cam, p)
        ctx.build_index()
        ctx.shared_load()
        hp = ctx.export_hitpoints()
        np.testing.assert_array_equal(ctx.export_stats(nh)["r2"], r2[hp["pix"][:, 0], hp["pix"][:, 1]])   # (c)
        for k in range(rounds):
            ctx.photon_pass(0, onetime, p * rounds + k)
            ctx.update()
        st = ctx.export_stats(nh)
        if p == 0:
            film_ref = ctx.resolve(rounds, onetime, w, h)
        ctx.shared_fold(rounds * onetime)
        emitted += rounds * onetime
        r2_before = r2.copy()
        oracle.shared_fold(r2, n, tau, hp["pix"], st["newN"], st["tao"], prm.alpha)
        assert np.all(r2 <= r2_before)
        got = ctx.shared_export()
        assert got["photons"] == emitted
        np.testing.assert_allclose(got["r2"], r2, rtol=1e-13)
        np.testing.assert_allclose(got["n"], n, rtol=1e-13)
        np.testing.assert_allclose(got["tau"], tau, rtol=1e-11, atol=1e-300)
        film = ctx.shared_resolve()
        if p == 0:
            np.testing.assert_allclose(film, film_ref, rtol=1e-12, atol=1e-300)        # (a)
        np.testing.assert_allclose(film, tau / (3.1415926 * r2[..., None] * emitted), rtol=1e-11, atol=1e-300)
    assert (n > 0).mean() > 0.5 and r2.min() < prm.init_r2 * prm.alpha


def test_jpeg_decode_encode_on_device(ctx):
    """nvJPEG replaces cv::imread (Element.h:47-59) / cv::imwrite (Raytracer.h:474).
    The CPU check is OpenCV's own decoder (the library family the reference uses): interleaved
    BGR, same size, pixel values equal up to the decoders' IDCT / upsampling rounding."""
    cv2 = pytest.importorskip("cv2")
    rng = np.random.RandomState(5)
    yy, xx = np.mgrid[0:203, 0:317]
    img = np.stack([127 + 100 * np.sin(xx / 23.0) * np.cos(yy / 17.0), (xx * 255.0 / 316), 255 - (yy * 255.0 / 202)], axis=2)
    img = np.clip(img + rng.normal(0, 2, img.shape), 0, 255).astype(np.uint8)
    # measured on B200 against OpenCV 4.13: 4:4:4 max 4 / mean 0.51; default 4:2:0 (different chroma
    # upsampling filters) max 6 / mean 0.99
    for flags, max_tol, mean_tol in (([cv2.IMWRITE_JPEG_QUALITY, 95, cv2.IMWRITE_JPEG_SAMPLING_FACTOR, cv2.IMWRITE_JPEG_SAMPLING_FACTOR_444], 6, 0.8),
                                     ([cv2.IMWRITE_JPEG_QUALITY, 90], 24, 1.5)):
        ok, enc = cv2.imencode(".jpg", img, flags)
        assert ok
        ref = cv2.imdecode(enc, cv2.IMREAD_COLOR)
        got = ctx.jpeg_decode(enc.tobytes())
        assert got.shape == ref.shape == img.shape and got.dtype == np.uint8
        d = np.abs(got.astype(int) - ref.astype(int))
        assert d.max() <= max_tol and d.mean() < mean_tol, (d.max(), d.mean())
    out = ctx.jpeg_encode(img, quality=95)
    assert out[:2] == b"\xff\xd8" and out[-2:] == b"\xff\xd9"
    back = cv2.imdecode(np.frombuffer(out, np.uint8), cv2.IMREAD_COLOR)
    assert back.shape == img.shape
    mse = np.mean((back.astype(float) - img.astype(float)) ** 2)
    assert 10 * np.log10(255.0 ** 2 / mse) > 35.0
    with pytest.raises(Exception):
        ctx.jpeg_decode(b"not a jpeg at all")


def test_eye_pass_row_sharding(ctx):
    """the eye pass restricted to row ranges, appended, is the eye pass of the whole
    film (draws are addressed by pixel): same hit points in canonical order, bit for bit."""
    sc = make_scene("reference", textures=None)
    cam = reference_camera(96, 72)
    ctx.set_params(_params())
    ctx.upload_scene(sc)
    n = ctx.eye_pass(cam, 3)
    whole = ctx.export_hitpoints()
    m = ctx.eye_pass_rows(cam, 3, 40, 72)
    assert 0 < m < n
    assert ctx.eye_pass_rows(cam, 3, 0, 17, append=True) > m
    k = ctx.eye_pass_rows(cam, 3, 17, 17, append=True)          # empty range: nothing added
    assert ctx.eye_pass_rows(cam, 3, 17, 40, append=True) == n > k
    parts = ctx.export_hitpoints()
    assert ctx.allgather_hitpoints() == parts["pos"].shape[0]   # no communicator: a no-op
    for k in ("pos", "n", "wgt", "r2", "pix", "node"):
        np.testing.assert_array_equal(parts[k], whole[k])
    with pytest.raises(Exception):
        ctx.eye_pass_rows(cam, 3, 10, 100)


def test_gpu_matches_the_reference_itself(ctx, oracle):
    """CUDA NearCollide against the reference's OWN compiled code (oracle/_ref, built from
    /root/reference/raytracer/*.h; the oracle only supplies the draw log that tells which
    Newton seeds the addressed RNG hands to each (ray, patch, restart))."""
    ref_mod = pytest.importorskip("oracle.ref")
    if not ref_mod.available():
        pytest.skip("oracle/_ref/libref.so not built (needs /root/reference)")
    ref = ref_mod.Reference()
    sc = make_scene("reference", textures="synthetic")
    prm = _params()
    ctx.set_params(prm)
    ctx.upload_scene(sc)
    rays = np.concatenate([rays_in_room(150, 41), rays_at_teapot(150, 42, sc)])
    ids = np.arange(rays.shape[0], dtype=np.uint64) + 11
    got = ctx.intersect_batch(rays, ids, pass_id=3, node=1)
    h = ref.scene(sc)
    bez = 0
    for i in range(rays.shape[0]):
        oracle.log_begin()
        oracle.intersect_batch(sc, prm, rays[i:i + 1], ids[i:i + 1], pass_id=3, node=1)
        vals, kinds = oracle.log_end()
        b = ref.near_collide(h, rays[i], vals, kinds)
        assert got["obj"][i] == b["obj"] and got["inside"][i] == b["inside"]
        assert got["t"][i] == b["t"] and np.array_equal(got["pos"][i], b["pos"])
        np.testing.assert_allclose(got["n"][i], b["n"], rtol=0, atol=1e-15)
        np.testing.assert_allclose(got["color"][i], b["color"], rtol=1e-12, atol=1e-15)
        bez += got["patch"][i] >= 0
    assert bez > 20


def test_full_size_properties(ctx):
    """BASELINE config 3 at full size (1024x768, 2 M photons): size-independent properties.
    (a) sum(newN) == accepted-pair counter and every pair was distance-tested first;
    (b) linearity: two half rounds == one full round (newN bit-exact, tao to rounding);
    (c) the film is the density estimate of tao (resolve), row-major [h][w][3];
    (d) per-photon statistics match the CPU probe (5.9 queries, 3.7 deposits);
    (e) FILTERED gather == EXACT gather: newN and the pair counters bit for bit."""
    sc = make_scene("reference", textures="synthetic")
    prm = _params(newton_mode=abi.SPPM_NEWTON_FAST)
    ctx.set_params(prm)
    ctx.upload_scene(sc)
    w, h, onetime = 1024, 768, 2_000_000
    cam = reference_camera(w, h)
    n = ctx.eye_pass(cam, 0)
    assert w * h - 16 <= n <= w * h  # 1.000 hit points per pixel in this scene (a few eye paths escape)
    ctx.build_index()
    ctx.reset_counters()
    ctx.photon_pass(0, onetime, 1)
    ctx.update()
    full = ctx.export_stats(n)
    c = ctx.counters()
    assert int(full["newN"].astype(np.int64).sum()) == c["acc_pairs"] > 0
    assert c["cand_pairs"] >= c["acc_pairs"] > 0.95 * c["cand_pairs"]
    assert c["photons"] == onetime
    assert 5.7 < c["isect_queries"] / onetime < 6.1 and 3.6 < c["deposits"] / onetime < 3.9
    assert c["isect_queries"] <= 15 * onetime  # Q9
    assert 0.03 < c["patch_solves"] / onetime < 0.045
    film = ctx.resolve(1, onetime, w, h)
    hp = ctx.export_hitpoints()
    want = np.zeros((h, w, 3))
    np.add.at(want, (hp["pix"][:, 0], hp["pix"][:, 1]), full["tao"] * (1.0 / 3.1415926 / (full["r2"] * 1 * onetime))[:, None])
    np.testing.assert_allclose(film, want, rtol=1e-12, atol=0)
    ctx.build_index()  # fresh accumulators, same hit points
    ctx.photon_pass(0, onetime // 2, 1)
    ctx.photon_pass(onetime // 2, onetime - onetime // 2, 1)
    ctx.update()
    halves = ctx.export_stats(n)
    np.testing.assert_array_equal(halves["newN"], full["newN"])
    np.testing.assert_allclose(halves["tao"], full["tao"], rtol=1e-11, atol=0)
    # (e) the packed-FP32 filtered gather (default) against the FP64-on-every-pair kernel on the
    # same 2 M photons: identical decisions on all 213 M candidate pairs
    ctx.set_params(_params(newton_mode=abi.SPPM_NEWTON_FAST, gather_mode=abi.SPPM_GATHER_EXACT))
    ctx.build_index()
    ctx.reset_counters()
    ctx.photon_pass(0, onetime, 1)
    ctx.update()
    exact = ctx.export_stats(n)
    ce = ctx.counters()
    np.testing.assert_array_equal(exact["newN"], full["newN"])
    np.testing.assert_allclose(exact["tao"], full["tao"], rtol=1e-11, atol=0)
    assert (ce["cand_pairs"], ce["acc_pairs"], ce["deposits_kept"]) == (c["cand_pairs"], c["acc_pairs"], c["deposits_kept"])


def test_newton_fast_sweep_against_exact(ctx):
    """40 000 (ray, patch) pairs: FAST Newton vs EXACT (bit-identical to the reference) — same
    hit/miss on every pair, (t, u, v) within the north-star 1e-6 relative."""
    sc = make_scene("reference", textures=None)
    n = 40_000
    patch = (np.arange(n) % 32).astype(np.int32)
    rays = rays_at_patches(patch, 77, sc)
    rays[: n // 4] = rays_at_teapot(n // 4, 78, sc)
    ctx.upload_scene(sc)
    out = {}
    for mode in (abi.SPPM_NEWTON_EXACT, abi.SPPM_NEWTON_FAST):
        ctx.set_params(_params(newton_mode=mode))
        out[mode] = ctx.bezier_patch_batch(rays, patch, None)
    ref, got = out[abi.SPPM_NEWTON_EXACT], out[abi.SPPM_NEWTON_FAST]
    hr, hg = ref[:, 0] < 1e19, got[:, 0] < 1e19
    assert hr.sum() > 20_000 and (~hr).sum() > 2_000
    assert (hr == hg).all()
    rel = np.abs(got[hr] - ref[hr]) / np.maximum(np.abs(ref[hr]), 1e-12)
    assert rel.max() <= 1e-6, rel.max()


# ---- against golden vectors produced by the reference's OWN code (tests/golden/ref_*.npz, frozen by
# tools/make_ref_golden.py from the reference's headers compiled against the OpenCV/FLANN shim) --------
def _ref_golden(name, scene=None):
    import os
    if scene == "quirks":   # the quirks scene's vectors use the same keys, in files of their own
        name = name.replace("ref_", "ref_quirks_")
    return np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", name + ".npz"))


def _ref_scene(name):
    from conftest import custom_scene, quirks_scene
    if name == "custom":
        return custom_scene()
    if name == "quirks":
        return quirks_scene()
    return make_scene(name, textures="synthetic" if name == "reference" else None)


def test_ref_golden_newton_exact(ctx):
    """Parity level 1 against the reference itself: Bezier3::newton with the reference's initial
    guesses injected — EXACT mode returns the reference's (t, u, v) bit for bit, FAST within 1e-6."""
    g = _ref_golden("ref_newton")
    sc = make_scene("reference", textures=None)
    ctx.set_params(_params(newton_mode=abi.SPPM_NEWTON_EXACT))
    ctx.upload_scene(sc)
    got = ctx.bezier_patch_batch(g["rays"], g["patch"], g["seeds"])
    np.testing.assert_array_equal(got, g["tuv"])
    assert 30 < (g["tuv"][:, 0] < 1e19).sum() < g["tuv"].shape[0]
    ctx.set_params(_params(newton_mode=abi.SPPM_NEWTON_FAST))
    fast = ctx.bezier_patch_batch(g["rays"], g["patch"], g["seeds"])
    hit = g["tuv"][:, 0] < 1e19
    assert ((fast[:, 0] < 1e19) == hit).all()
    rel = np.abs(fast[hit] - g["tuv"][hit]) / np.maximum(np.abs(g["tuv"][hit]), 1e-12)
    assert rel.max() <= 1e-6  # north-star tolerance on (t, u, v)


@pytest.mark.parametrize("name", ["cornell", "custom", "reference", "stress"])
def test_ref_golden_near_collide(ctx, name):
    """Scene::NearCollide as computed by the reference's own PlaneObj / SphereObj / Bezier3Obj classes."""
    g = _ref_golden("ref_near_collide", name)
    ctx.set_params(_params())
    ctx.upload_scene(_ref_scene(name))
    rays = g[name + "_rays"]
    got = ctx.intersect_batch(rays, np.arange(rays.shape[0], dtype=np.uint64), pass_id=1, node=2)
    np.testing.assert_array_equal(got["obj"], g[name + "_obj"])
    np.testing.assert_array_equal(got["inside"], g[name + "_inside"])
    np.testing.assert_array_equal(got["t"], g[name + "_t"])      # + - * / sqrt only: bit-exact
    np.testing.assert_array_equal(got["pos"], g[name + "_pos"])
    np.testing.assert_allclose(got["n"], g[name + "_n"], rtol=0, atol=1e-15)
    np.testing.assert_allclose(got["color"], g[name + "_color"], rtol=1e-12, atol=1e-15)  # sphere texture: acos / sin


@pytest.mark.parametrize("name", ["custom", "reference"])
def test_ref_golden_view_trace(ctx, name):
    """GetHitPoint / ViewTrace: the hit points the reference's own recursion produces for a grid of pixels
    (mirror and glass chains, branching materials, the teapot) are the eye pass's hit points of those pixels."""
    g = _ref_golden("ref_view_trace", name)
    w, h = (int(v) for v in g["film"])
    ctx.set_params(_params())
    ctx.upload_scene(_ref_scene(name))
    ctx.eye_pass(reference_camera(w, h), pass_id=0)
    got = ctx.export_hitpoints()
    gp = got["pix"][:, 0].astype(np.int64) * w + got["pix"][:, 1]
    off = 0
    for (i, j), cnt in zip(g[name + "_pixels"], g[name + "_counts"]):
        sel = np.nonzero(gp == int(i) * w + int(j))[0]
        assert sel.size == cnt, (i, j)
        want = {f: g[f"{name}_{f}"][off:off + cnt] for f in ("pos", "n", "wgt", "r2")}
        off += cnt
        # the reference appends in recursion order, the GPU exports in (pixel, path node) order: compare as sets
        a = sel[np.lexsort(got["pos"][sel].T[::-1])]
        b = np.lexsort(want["pos"].T[::-1])
        np.testing.assert_array_equal(got["pos"][a], want["pos"][b])
        np.testing.assert_allclose(got["n"][a], want["n"][b], rtol=0, atol=1e-15)
        np.testing.assert_allclose(got["wgt"][a], want["wgt"][b], rtol=1e-12, atol=0)
        np.testing.assert_array_equal(got["r2"][a], want["r2"][b])
    assert off == g[name + "_pos"].shape[0]


@pytest.mark.parametrize("name", ["cornell", "custom", "reference"])
def test_ref_golden_photon_pass(ctx, name):
    """Light::emit + PhotonTrace + radius search / accumulate as run by the reference's own code, 300
    photons per light: the GPU photon pass over the same hit points gives the same newN (bit-exact
    except where ulp differences of acos / sin / cos between libm and CUDA move a deposit across a
    radius boundary or flip a path decision: near-total agreement required) and tao within 1e-5."""
    g = _ref_golden("ref_photon", name)
    hp = {k: g[f"{name}_hp_{k}"] for k in ("pos", "n", "wgt", "r2", "pix")}
    ctx.set_params(_params())
    ctx.upload_scene(_ref_scene(name))
    ctx.import_hitpoints(hp["pos"], hp["n"], hp["wgt"], hp["r2"], hp["pix"], 64, 48)
    ctx.build_index()
    count = g[f"{name}_tot_newN"].shape[0]
    ctx.photon_pass(0, count, 5)
    ctx.update()
    got = ctx.export_stats(hp["pos"].shape[0])
    want_n, want_t = g[f"{name}_acc_newN"], g[f"{name}_acc_tao"]
    assert want_n.sum() > 50
    assert (got["newN"] == want_n).mean() > 0.995
    assert abs(int(got["newN"].sum()) - int(want_n.sum())) <= max(2, 0.01 * want_n.sum())
    same = got["newN"] == want_n
    np.testing.assert_allclose(got["tao"][same], want_t[same], rtol=1e-5, atol=1e-12)
