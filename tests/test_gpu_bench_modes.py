"""End-to-end parity of the configuration bench.py times — FAST Newton (monomial Horner + FMA), FAST
sampling math, bounce-synchronous wavefront tracer, FP32-filtered gather (bench.make_params + the
--trace / --gather defaults) — against the CPU oracle, which follows the reference operation by
operation (Bezier.h:112-159, Vec3.h:90-115, Raytracer.h:117-209, 296-357).

The pieces are held one by one in test_gpu_parity.py (FAST Newton <= 1e-6 per (ray, patch) pair, FAST
directions vs EXACT ones, filtered gather vs the exhaustive search); here their composition is:
eye pass, photon pass (deposit list per (photon, bounce), gathered newN / tao) and film, on the scenes of
BASELINE configs 1 (as specified: 320x240, 100 k photons), 2 / 4 (Cornell object set) and 5 (four
teapots) at photon counts the single-threaded oracle finishes in seconds.

Bars: positions <= 1e-6 relative (north-star level 1); the same deposit for >= 99.95 % of the (photon,
bounce) pairs (CUDA's sqrt-free transcendentals differ from libm's by ulps, which can flip a discrete
decision on a rare path and re-route that photon); newN equal on >= 99.95 % of the hit points with totals
within 0.02 %; film rRMSE <= 1 % (level 3).  Measured on the B200 (tools/test_margins.py,
profiles/r2_margins.txt): every deposit matched (positions <= 1.5e-9), newN equal on every hit point, the
counters identical, tao sums to 1e-15 — the bars leave room for another libm, not for a different algorithm."""
import numpy as np
import pytest

import bench
from raytrace3_b200 import RayTracer, abi
from raytrace3_b200.scene import make_scene, reference_camera

pytestmark = pytest.mark.gpu


def bench_params():
    """Exactly what bench.py's default arm runs."""
    p = bench.make_params(abi, fast=True)
    p.trace_mode = abi.SPPM_TRACE_WAVEFRONT
    p.gather_mode = abi.SPPM_GATHER_FILTERED
    p.count_pairs = 0       # the timed loop runs gather_filtered_kernel<COUNT = false>
    return p


def exact_params():
    return abi.default_params()


def _scene(name):
    return make_scene(name, textures=None if name == "cornell" else "synthetic")


def _match(keys_a, keys_b):
    """Indices (ia, ib) of the keys present in both (keys unique within each array)."""
    _, ia, ib = np.intersect1d(keys_a, keys_b, assume_unique=True, return_indices=True)
    return ia, ib


def _eye_check(ctx, oracle, scene_name, w, h, min_common=0.9999):
    sc = _scene(scene_name)
    ctx.set_params(bench_params())
    ctx.upload_scene(sc)
    cam = reference_camera(w, h)
    ref = oracle.eye_pass(sc, exact_params(), cam, pass_id=0)
    n = ctx.eye_pass(cam, pass_id=0)
    got = ctx.export_hitpoints()
    kr = (ref["pix"][:, 0].astype(np.int64) * w + ref["pix"][:, 1]) * (1 << 20) + ref["node"]
    kg = (got["pix"][:, 0].astype(np.int64) * w + got["pix"][:, 1]) * (1 << 20) + got["node"]
    ig, ir = _match(kg, kr)
    # the same (pixel, eye-path node) set: FAST Newton may flip hit / miss on a grazing ray, nothing else may
    assert len(ig) >= min_common * max(n, len(kr)), (n, len(kr), len(ig))
    rel = np.abs(got["pos"][ig] - ref["pos"][ir]) / np.maximum(1.0, np.abs(ref["pos"][ir]))
    assert rel.max() <= 1e-6, rel.max()                                     # north-star level 1
    assert np.abs(got["n"][ig] - ref["n"][ir]).max() <= 1e-6
    np.testing.assert_allclose(got["wgt"][ig], ref["wgt"][ir], rtol=1e-9, atol=0)
    return sc, cam, got


def test_bench_modes_eye_pass_config1(ctx, oracle):
    """BASELINE config 1 as specified: main.cpp scene at 320x240 — every hit point of the FAST eye pass within
    1e-6 of the reference's (Raytracer.h:296-357 over Bezier.h:112-159)."""
    _eye_check(ctx, oracle, "reference", 320, 240)


def test_bench_modes_eye_pass_four_teapots(ctx, oracle):
    """BASELINE config 5's scene (four Bezier objects) at reduced film size."""
    _eye_check(ctx, oracle, "stress", 384, 216)


def _photon_check(ctx, oracle, scene_name, w, h, count, pass_id, begin=0):
    sc, cam, _ = _eye_check(ctx, oracle, scene_name, w, h)
    prm = bench_params()
    ctx.build_index()
    hp = ctx.export_hitpoints()
    ref = oracle.photon_pass(sc, exact_params(), begin, count, pass_id, hp=hp, keep_deposits=True)
    rc = ref["counters"]
    # (a) the deposit list, per (photon, bounce)
    ctx.keep_deposits(True)
    try:
        ctx.reset_counters()
        ctx.photon_pass(begin, count, pass_id)
        ctx.update()
        dep = ctx.export_deposits()
        c = ctx.counters()
    finally:
        ctx.keep_deposits(False)
    assert c["photons"] == rc["photons"] == count * len(sc.plvec)
    assert abs(c["isect_queries"] - rc["queries"]) <= 0.0002 * rc["queries"]
    assert abs(c["deposits"] - rc["deposits"]) <= 0.0002 * rc["deposits"]
    assert abs(c["patch_solves"] - rc["solves"]) <= 0.001 * rc["solves"] + 2
    rd = ref["dep"]
    kg = dep["photon"].astype(np.int64) * 64 + dep["bounce"]
    kr = rd["photon"].astype(np.int64) * 64 + rd["bounce"]
    ig, ir = _match(kg, kr)
    rel = (np.abs(dep["x"][ig] - rd["x"][ir]) / np.maximum(1.0, np.abs(rd["x"][ir]))).max(1)
    same = rel <= 1e-6
    assert same.sum() >= 0.9995 * len(kr), (int(same.sum()), len(kr), len(kg))
    # flux = product of the surface colours met so far: a bilinear texture lookup turns a 1e-7 difference in the
    # hit position into a ~1e-4 difference in colour, so the bar is 1e-6 for nearly all and 2e-3 for every deposit
    fg, fr = dep["flux"][ig][same], rd["flux"][ir][same]
    assert np.isclose(fg, fr, rtol=1e-6, atol=1e-12).mean() >= 0.9995
    np.testing.assert_allclose(fg, fr, rtol=2e-3, atol=1e-12)
    assert np.abs(dep["n"][ig][same] - rd["n"][ir][same]).max() <= 1e-6
    # (b) the pass as the bench runs it (deposits outside the hit points' reach dropped, lanes, binning,
    # filtered gather): accumulators of a fresh map against the oracle's kd-tree gather
    ctx.build_index()          # fresh accumulators
    ctx.reset_counters()
    ctx.photon_pass(begin, count, pass_id)
    ctx.update()
    got = ctx.export_stats(hp["pos"].shape[0])
    assert ref["newN"].sum() > 0
    assert (got["newN"] == ref["newN"]).mean() >= 0.9995
    assert abs(int(got["newN"].sum()) - int(ref["newN"].sum())) <= 0.0002 * ref["newN"].sum()
    np.testing.assert_allclose(got["tao"].sum(0), ref["tao"].sum(0), rtol=2e-4)
    # level 2 on the identical photon list: the GPU's own deposits through the oracle's exact kd-tree search
    # (held equal to the exhaustive scan by tests/test_oracle_cpu.py::test_kdtree_equals_brute)
    newN, tao = oracle.gather_kdtree(hp, dep, prm.init_r2)
    np.testing.assert_array_equal(got["newN"], newN)
    np.testing.assert_allclose(got["tao"], tao, rtol=1e-5, atol=1e-300)
    return sc, cam


def test_bench_modes_photon_pass_config1(ctx, oracle):
    """BASELINE config 1 as specified: 320x240, one photon pass of 100 000 photons."""
    _photon_check(ctx, oracle, "reference", 320, 240, 100_000, 1)


def test_bench_modes_photon_pass_cornell(ctx, oracle):
    """BASELINE configs 2 / 4 (Cornell object set: mirror + glass spheres, no Bezier) at 256x192, 60 k photons."""
    _photon_check(ctx, oracle, "cornell", 256, 192, 60_000, 2, begin=1000)


def test_bench_modes_photon_pass_four_teapots(ctx, oracle):
    """BASELINE config 5's scene at 192x108, 40 k photons."""
    _photon_check(ctx, oracle, "stress", 192, 108, 40_000, 3)


@pytest.mark.parametrize("scene_name,w,h,cnt,onetime", [("reference", 64, 48, 2, 30000), ("cornell", 96, 54, 2, 30000)])
def test_bench_modes_image_rrmse(ctx, oracle, scene_name, w, h, cnt, onetime):
    """Level 3: the film RayTracer::render() makes in the bench's modes vs the oracle's, matched photon counts."""
    sc = _scene(scene_name)
    ctx.set_params(bench_params())
    cam = reference_camera(w, h)
    rt = RayTracer(ctx)
    rt.camera, rt.scene, rt.cnt, rt.onetime = cam, sc, cnt, onetime
    rt.render()
    ref = oracle.render_replicas(sc, exact_params(), cam.pos, cam.look, w, h, cnt, onetime, threads=1, jitter=0.0)
    film_ref, film = ref["film"], cam.pic
    assert film_ref.max() > 0
    rrmse = np.sqrt(((film - film_ref) ** 2).mean()) / np.sqrt((film_ref ** 2).mean())
    assert rrmse <= 0.01, rrmse


def test_bench_modes_progressive_rounds(ctx, oracle):
    """BASELINE config 5's update rule (radius-shrinking passes, PROGRESSIVE) in the bench's modes: three rounds
    of trace + gather + HitPoint::update against the oracle's trace + kd-tree gather + update (Raytracer.h:69-79)."""
    sc, cam, _ = _eye_check(ctx, oracle, "stress", 160, 90)
    prm = bench_params()
    prm.update_mode = abi.SPPM_UPDATE_PROGRESSIVE
    ctx.set_params(prm)
    ctx.build_index()
    hp = ctx.export_hitpoints()
    n = hp["pos"].shape[0]
    eprm = exact_params()
    eprm.update_mode = abi.SPPM_UPDATE_PROGRESSIVE
    N = np.zeros(n, dtype=np.int32)
    newN = np.zeros(n, dtype=np.int32)
    r2 = hp["r2"].copy()
    tao = np.zeros((n, 3))
    for rnd in range(3):
        ctx.photon_pass(0, 30000, 10 + rnd)
        ctx.update()
        hpr = dict(hp)
        hpr["r2"] = r2
        ref = oracle.photon_pass(sc, eprm, 0, 30000, 10 + rnd, hp=hpr, newN=newN, tao=tao)
        oracle.update(abi.SPPM_UPDATE_PROGRESSIVE, prm.alpha, newN, N, r2, tao)
    got = ctx.export_stats(n)
    assert (got["N"] == N).mean() >= 0.99
    assert abs(int(got["N"].sum()) - int(N.sum())) <= 0.003 * N.sum()
    same = got["N"] == N
    np.testing.assert_allclose(got["r2"][same], r2[same], rtol=1e-9)
    assert r2.min() < 0.9 * prm.init_r2          # the radii did shrink
    np.testing.assert_allclose(got["tao"].sum(0), tao.sum(0), rtol=3e-3)


def _light_over_teapot_scene():
    """The light just above the teapot's bounding box: every downward ray enters the object box and several
    patch boxes (1.5 Newton solves per emitted photon; the shipped scene makes 0.037)."""
    sc = make_scene("reference", textures=None)
    bz = [o for o in sc.objvec if o.kind == 2][0]
    lo, hi = bz.P.reshape(-1, 3).min(0), bz.P.reshape(-1, 3).max(0)
    sc.plvec[0].pos = tuple(0.5 * (lo + hi) + np.array([0.0, 0.6 * (hi[1] - lo[1]), 0.0]))
    return sc


@pytest.mark.parametrize("scene_name,item_cap", [("stress", 256), ("light_over_teapot", 4096), ("light_over_teapot", 0)])
def test_wavefront_item_list_spill(ctx, monkeypatch, scene_name, item_cap):
    """The wavefront tracer's (ray, patch) item list is sized for 2 items per photon and bounce.  A bounce that
    finds it full must not drop Bezier solves (Bezier.h:240-282 runs for every ray that enters a patch box): those
    photons are finished by the persistent tracer.  Deposits, gathered statistics and counters equal the persistent
    tracer's bit for bit, with the list cut short (SPPM_WF_ITEM_CAP) and at the product's sizing."""
    sc = _light_over_teapot_scene() if scene_name == "light_over_teapot" else make_scene(scene_name, textures=None)
    cam = reference_camera(64, 48)
    out = []
    for mode in (abi.SPPM_TRACE_WAVEFRONT, abi.SPPM_TRACE_PERSISTENT):
        prm = bench_params()
        prm.trace_mode = mode
        ctx.set_params(prm)
        ctx.upload_scene(sc)
        n = ctx.eye_pass(cam, 0)
        ctx.build_index()
        ctx.keep_deposits(True)
        if item_cap and mode == abi.SPPM_TRACE_WAVEFRONT:
            monkeypatch.setenv("SPPM_WF_ITEM_CAP", str(item_cap))
        try:
            ctx.reset_counters()
            ctx.photon_pass(3, 40000, 6)
            ctx.update()
            out.append((ctx.export_deposits(), ctx.export_stats(n), ctx.counters()))
        finally:
            ctx.keep_deposits(False)
            monkeypatch.delenv("SPPM_WF_ITEM_CAP", raising=False)
    (d0, s0, c0), (d1, s1, c1) = out
    for k in ("photon", "bounce", "x", "n", "flux"):
        np.testing.assert_array_equal(d0[k], d1[k], err_msg=k)
    np.testing.assert_array_equal(s0["newN"], s1["newN"])
    np.testing.assert_allclose(s0["tao"], s1["tao"], rtol=1e-12)
    for k in ("photons", "isect_queries", "patch_solves", "deposits", "deposits_kept", "acc_pairs"):
        assert c0[k] == c1[k], k
    assert c0["patch_solves"] > 1000
    if item_cap:
        assert c0["spilled_photons"] > 0
    assert c1["spilled_photons"] == 0


def test_texture_levels_on_device(ctx, oracle):
    """textures are stored in 8 x 4 tiles with a 2 x 2 box-filtered mip chain below them.  tex_lod = 0
    (default) is Texture::colorUV on the uploaded image, bit for bit (the textured NearCollide / eye-pass parity tests
    run on it); tex_lod = 2 samples the second level: the same hit points, colours close to the full-resolution ones
    and equal to the oracle's sampler on the box-filtered image."""
    sc = make_scene("reference", textures="synthetic")
    cam = reference_camera(160, 120)
    out = []
    for lod in (0, 2):
        prm = bench_params()
        prm.tex_lod = lod
        ctx.set_params(prm)
        ctx.upload_scene(sc)
        n = ctx.eye_pass(cam, 0)
        out.append(ctx.export_hitpoints())
    ctx.set_params(bench_params())
    a, b = out
    np.testing.assert_array_equal(a["pos"], b["pos"])          # geometry does not depend on the level
    np.testing.assert_array_equal(a["pix"], b["pix"])
    textured = np.abs(a["wgt"] - b["wgt"]).max(1) > 0
    assert 0.2 < textured.mean() < 1.0                          # the textured walls / spheres changed, the rest did not
    assert np.abs(a["wgt"] - b["wgt"]).max() < 0.5 and np.abs(a["wgt"] - b["wgt"]).mean() < 0.05
    ref = oracle.eye_pass(sc, exact_params(), cam, pass_id=0)
    order = np.lexsort((ref["node"], ref["pix"][:, 0] * cam.w + ref["pix"][:, 1]))
    np.testing.assert_allclose(a["wgt"], ref["wgt"][order], rtol=1e-9, atol=0)   # level 0 == the reference's sampler


@pytest.mark.parametrize("workload,rounds", [("caustics-1080p", 1), ("stress-4k", 3)])
def test_full_size_dense_maps_filtered_equals_exact(ctx, workload, rounds):
    """BASELINE configs 4 / 5 at their full film sizes (1920x1080: 2.07 M hit points; 3840x2160 with four teapots:
    8.29 M) — the dense maps on which a work unit is small against the radius, so the filtered gather's whole-unit
    accept, Morton-ordered units and (config 5, PROGRESSIVE) per-unit cull radius all act.  Size-independent properties
    over `rounds` rounds of 2 M photons in the bench's modes:
    (a) the FP32-filtered gather (count_pairs = 0: the timed build) == the FP64-on-every-pair kernel on the same
        photons: N, newN and R2 bit for bit, tao to rounding;
    (b) sum of accepted pairs == the accept counter;
    (c) PROGRESSIVE: the radii only shrink, and do so wherever photons were counted."""
    wl = bench.WORKLOADS[workload]
    sc = make_scene(wl["scene"])
    cam = reference_camera(wl["w"], wl["h"])
    progressive = workload == "stress-4k"
    onetime = 2_000_000
    out = []
    for gmode in (abi.SPPM_GATHER_FILTERED, abi.SPPM_GATHER_EXACT):
        prm = bench_params()
        prm.gather_mode = gmode
        prm.update_mode = abi.SPPM_UPDATE_PROGRESSIVE if progressive else abi.SPPM_UPDATE_REFERENCE
        ctx.set_params(prm)
        ctx.upload_scene(sc)
        n = ctx.eye_pass(cam, 0)
        ctx.build_index()
        ctx.reset_counters()
        for r in range(rounds):
            ctx.photon_pass(0, onetime, 1 + r)
            ctx.update()
        out.append((ctx.export_stats(n), ctx.counters()))
    ctx.set_params(bench_params())
    (f, cf), (e, ce) = out
    assert n > 0.99 * wl["w"] * wl["h"]
    np.testing.assert_array_equal(f["newN"], e["newN"])
    np.testing.assert_array_equal(f["N"], e["N"])
    np.testing.assert_array_equal(f["r2"], e["r2"])
    np.testing.assert_allclose(f["tao"], e["tao"], rtol=1e-10, atol=1e-300)
    assert cf["acc_pairs"] == ce["acc_pairs"] > n
    assert cf["deposits_kept"] == ce["deposits_kept"]
    if progressive:
        got_photons = f["N"] > 0          # (a single photon leaves N = int(0.7) = 0 but shrinks the radius all the same)
        assert got_photons.mean() > 0.5
        assert (f["r2"] <= 2.0).all() and (f["r2"][got_photons] < 2.0).all() and (f["r2"] == 2.0).any()
    else:
        assert int(f["newN"].astype(np.int64).sum()) == cf["acc_pairs"]
