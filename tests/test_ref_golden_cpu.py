"""The oracle against golden vectors produced by THE REFERENCE'S OWN CODE (tests/golden/ref_*.npz,
frozen by tools/make_ref_golden.py from oracle/_ref/libref.so = the reference's headers compiled
against the OpenCV/FLANN shim).  Unlike tests/test_ref_pin_cpu.py this needs neither
/root/reference nor the git-ignored libref.so: the reference's outputs travel as fixtures.
Every comparison is bit-exact (same arithmetic, same host libm)."""
import os

import numpy as np
import pytest

from conftest import custom_scene, quirks_scene
from raytrace3_b200 import abi
from raytrace3_b200.scene import make_scene, reference_camera, teapot_patches

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
SCENES = {
    "cornell": lambda: make_scene("cornell", textures=None),
    "reference": lambda: make_scene("reference", textures="synthetic"),
    "stress": lambda: make_scene("stress", textures=None),
    "custom": custom_scene,
    "quirks": quirks_scene,   # Q6 textured teapot, Q8 zero diffuse channels; its vectors live in ref_quirks_*.npz
}


def _load(name, scene=None):
    return np.load(os.path.join(GOLD, (name.replace("ref_", "ref_quirks_") if scene == "quirks" else name) + ".npz"))


def test_pieces(oracle):
    """evalPatchPoint/DerivU/DerivV, AABBbox::intersect, GetUnitRandRefl, refracted, Camera::lookAt/emit,
    Texture::colorUV, HitPoint::update, toInt, the resolve expression."""
    g = _load("ref_pieces")
    P = teapot_patches()
    for k in range(g["be_idx"].shape[0]):
        got = np.stack(oracle.bezier_eval(P[g["be_idx"][k]] * g["be_scale"][k], g["be_uv"][k, 0], g["be_uv"][k, 1]))
        assert np.array_equal(got, g["be_out"][k]), k
    hits = np.array([oracle.aabb_intersect(g["ab_box"][k], g["ab_org"][k], g["ab_dir"][k]) for k in range(g["ab_box"].shape[0])])
    assert np.array_equal(hits, g["ab_hit"]) and 100 < hits.sum() < hits.size - 100
    for k in range(g["sr_n"].shape[0]):
        assert np.array_equal(oracle.unit_rand_refl(g["sr_n"][k], g["sr_u"][k, 0], g["sr_u"][k, 1]), g["sr_out"][k]), k
        assert np.array_equal(oracle.refracted(g["rf_I"][k], g["rf_N"][k], 1.5), g["rf_out"][k, 0]), k
        assert np.array_equal(oracle.refracted(g["rf_I"][k], g["rf_N"][k], 1 / 1.5), g["rf_out"][k, 1]), k
    for c in g["cam"]:
        w, h, y, x = (int(v) for v in c[6:10])
        oc = oracle.camera_lookat(c[0:3], c[3:6], w, h)
        assert np.array_equal(np.concatenate([list(oc.pos), list(oc.dir), list(oc.du), list(oc.dv)]), c[10:22])
        assert np.array_equal(oracle.camera_emit(oc, y, x), c[22:28])
    for k in range(g["tx_uv"].shape[0]):
        assert np.array_equal(oracle.color_uv(g["tx_map"], g["tx_uv"][k, 0], g["tx_uv"][k, 1]), g["tx_out"][k]), k
    a = [g[f].copy() for f in ("up_newN", "up_N", "up_r2", "up_tao")]
    oracle.update(abi.SPPM_UPDATE_REFERENCE, 0.7, *a)
    for x, f in zip(a, ("up_o_newN", "up_o_N", "up_o_r2", "up_o_tao")):
        assert np.array_equal(x, g[f]), f
    x = g["ti_x"]
    img = oracle.tonemap(np.repeat(x, 3).reshape(1, -1, 3), 1.0)   # one film row, every value in all three channels
    assert np.array_equal(img[0, :, 0].astype(np.int32), g["ti_out"])
    for r, want in zip(g["rs_in"], g["rs_out"]):
        f = oracle.resolve(np.array([[0, 0]], np.int32), np.array([r[3]]), r[None, :3].copy(), 7, 100, 1, 1)
        assert np.array_equal(f[0, 0], want)


def test_newton(oracle):
    """Bezier3::newton incl. cv::Matx33d::inv (Bezier.h:112-159), 50 injected restarts."""
    g = _load("ref_newton")
    P = make_scene("reference", textures=None).objvec[8].P
    hits = 0
    for i in range(g["rays"].shape[0]):
        got = oracle.newton(P[g["patch"][i]], g["rays"][i, :3], g["rays"][i, 3:], g["seeds"][i])
        assert np.array_equal(got, g["tuv"][i]), i
        hits += got[0] < 1e19
    assert 30 < hits < g["rays"].shape[0]
    fs = np.full((50, 2), 0.5)
    for r, want in zip(g["flat_rays"], g["flat_tuv"]):
        assert np.array_equal(oracle.newton(g["flat_P"], r[:3], r[3:], fs), want)
    # and as one batch through the oracle's batch entry (what the GPU test compares against)
    sc = make_scene("reference", textures=None)
    got = oracle.bezier_patch_batch(sc, abi.default_params(), g["rays"], g["patch"], g["seeds"])
    assert np.array_equal(got, g["tuv"])


@pytest.mark.parametrize("name", sorted(SCENES))
def test_near_collide(oracle, params, name):
    """Scene::NearCollide through the reference's own PlaneObj / SphereObj / Bezier3Obj."""
    g = _load("ref_near_collide", name)
    sc = SCENES[name]()
    rays = g[name + "_rays"]
    got = oracle.intersect_batch(sc, params, rays, np.arange(rays.shape[0], dtype=np.uint64), pass_id=1, node=2)
    for f in ("obj", "inside", "t", "pos", "n", "color"):
        assert np.array_equal(got[f], g[f"{name}_{f}"]), f
    if name in ("reference", "stress", "quirks"):
        assert (got["patch"] >= 0).sum() > 15
    if name == "quirks":   # Q6: the teapot's colour comes from its texture, looked up at (u, t)
        tp = got["obj"] == 8
        assert len(np.unique(got["color"][tp].round(6), axis=0)) > 10


@pytest.mark.parametrize("name", ["cornell", "custom", "quirks", "reference"])
def test_photon_trace(oracle, params, name):
    """Light::emit + PhotonTrace + radiusSearch / accumulate: per-photon totals and the sum over 300 photons."""
    g = _load("ref_photon", name)
    sc = SCENES[name]()
    hp = {k: g[f"{name}_hp_{k}"] for k in ("pos", "n", "wgt", "r2", "pix")}
    eye = oracle.eye_pass(sc, params, reference_camera(64, 48), 0)   # the stored hit points are the oracle's eye pass
    for k in hp:
        assert np.array_equal(eye[k], hp[k]), k
    n = hp["pos"].shape[0]
    acc_n, acc_t = np.zeros(n, np.int64), np.zeros((n, 3))
    for j in range(g[f"{name}_tot_newN"].shape[0]):
        r = oracle.photon_pass(sc, params, j, 1, 5, hp=hp, use_kdtree=(j % 2 == 0))
        assert int(r["newN"].sum()) == g[f"{name}_tot_newN"][j], j
        assert np.array_equal(r["tao"].sum(axis=0), g[f"{name}_tot_tao"][j]), j
        acc_n += r["newN"]
        acc_t += r["tao"]
    assert np.array_equal(acc_n, g[f"{name}_acc_newN"]) and np.array_equal(acc_t, g[f"{name}_acc_tao"])
    assert acc_n.sum() > 50


@pytest.mark.parametrize("name", ["custom", "quirks", "reference"])
def test_view_trace(oracle, params, name):
    """ViewTrace: mirror / glass chains, branching materials, the teapot."""
    g = _load("ref_view_trace", name)
    w, h = (int(v) for v in g["film"])
    sc = SCENES[name]()
    cam = reference_camera(w, h)
    off = 0
    for (i, j), cnt in zip(g[name + "_pixels"], g[name + "_counts"]):
        a = oracle.view_trace_pixel(sc, params, cam, int(i), int(j))
        assert a["pos"].shape[0] == cnt, (i, j)
        for f in ("pos", "n", "wgt", "r2", "pix"):
            assert np.array_equal(a[f], g[f"{name}_{f}"][off:off + cnt]), (i, j, f)
        off += cnt
    assert off == g[name + "_pos"].shape[0]
    if name == "custom":
        assert off > len(g[name + "_counts"])  # branching produced extra hit points
    if name == "quirks":
        assert off < len(g[name + "_counts"])  # Q8: no hit point on an object with a zero diffuse channel
