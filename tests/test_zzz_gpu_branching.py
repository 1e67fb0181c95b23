"""GPU eye pass on a scene whose eye paths branch at most hits (conftest.branching_scene): more hit
points than the 4 per pixel sppm_eye_pass_rows reserves up front, so the pass is traced a second time
into a buffer of the size the first one counted.  Written after the round's GPU budget was spent (first
run by the round-end driver); the kernel itself is held to the oracle on this scene by the CPU
emulation (tests/test_device_host_cpu.py::test_eye_pass_kernel_branching_overflow)."""
import numpy as np
import pytest

from conftest import branching_scene
from raytrace3_b200 import abi
from raytrace3_b200.scene import reference_camera

pytestmark = pytest.mark.gpu


def test_eye_pass_grows_its_buffer(ctx, oracle):
    sc = branching_scene()
    prm = abi.default_params()
    ctx.set_params(prm)
    ctx.upload_scene(sc)
    w, h = 48, 36
    cam = reference_camera(w, h)
    ref = oracle.eye_pass(sc, prm, cam, pass_id=1)
    assert ref["pos"].shape[0] > 4 * w * h + 1024
    n = ctx.eye_pass(cam, pass_id=1)
    assert n == ref["pos"].shape[0]
    ctx.build_index()
    got = ctx.export_hitpoints()
    order = np.lexsort((ref["node"], ref["pix"][:, 0] * w + ref["pix"][:, 1]))
    np.testing.assert_array_equal(got["pix"], ref["pix"][order])
    np.testing.assert_array_equal(got["node"], ref["node"][order])
    np.testing.assert_array_equal(got["pos"], ref["pos"][order])
    np.testing.assert_allclose(got["wgt"], ref["wgt"][order], rtol=1e-12, atol=0)
    # and a smaller film afterwards still works (the buffer only ever grows)
    n2 = ctx.eye_pass(reference_camera(16, 12), pass_id=1)
    assert n2 == oracle.eye_pass(sc, prm, reference_camera(16, 12), pass_id=1)["pos"].shape[0]


def test_film_sums_a_pixels_hit_points_in_canonical_order(ctx, oracle):
    """Raytracer.h:281-293 on a map with many hit points per pixel: sppm_resolve == the oracle's film of the
    exported state (canonical order = pixel, eye-path node, the order GetHitPoint appends to Hvec) bit for
    bit, and a second resolve gives the same bits (no arrival-order atomics between a pixel's hit points)."""
    sc = branching_scene()
    prm = abi.default_params()
    ctx.set_params(prm)
    ctx.upload_scene(sc)
    w, h, onetime = 48, 36, 200_000
    n = ctx.eye_pass(reference_camera(w, h), pass_id=1)
    ctx.build_index()
    ctx.photon_pass(0, onetime, 1)
    ctx.update()
    hp = ctx.export_hitpoints()
    st = ctx.export_stats(n)
    assert n > 4 * w * h and st["tao"].sum() > 0
    want = oracle.resolve(hp["pix"], st["r2"], st["tao"], 1, onetime, w, h)
    film = ctx.resolve(1, onetime, w, h)
    np.testing.assert_array_equal(film, want)
    np.testing.assert_array_equal(ctx.resolve(1, onetime, w, h), film)


def test_resolve_into_a_page_locked_film(ctx):
    """sppm_host_pin / sppm_host_unpin (what the drop-in Camera does with camera.pic): the film read back into a
    page-locked caller buffer is the film read back into pageable memory; a null pointer is refused."""
    from raytrace3_b200 import api
    sc = branching_scene()
    ctx.set_params(abi.default_params())
    ctx.upload_scene(sc)
    w, h, onetime = 48, 36, 50_000
    ctx.eye_pass(reference_camera(w, h), pass_id=1)
    ctx.build_index()
    ctx.photon_pass(0, onetime, 1)
    ctx.update()
    plain = ctx.resolve(1, onetime, w, h)
    film = np.full((h, w, 3), -1.0)
    assert api.host_pin(film)
    try:
        ctx.resolve(1, onetime, w, h, out=film)
        np.testing.assert_array_equal(film, plain)
    finally:
        assert api.host_unpin(film)
    assert ctx.lib.sppm_host_pin(None, 16) == -1 and ctx.lib.sppm_host_unpin(None) == -1   # SPPM_ERR_INVALID
    assert not api.host_unpin(film)      # not registered any more: an error code, no crash
