"""GPU parity on the `quirks` scene (conftest.quirks_scene): branches of the reference the shipped scene
never reaches -- a textured teapot (colour looked up at (u, t), Bezier.h:278, Q6) and materials with a
zero diffuse channel (not IsDiff(), yet Roulette still takes the DIFF branch, Q8) -- against the golden
vectors the reference's own code produced for it (tests/golden/ref_quirks_*.npz,
tools/make_ref_golden.py).  Same checks and bars as test_ref_golden_* in test_gpu_parity.py."""
import numpy as np
import pytest

import test_gpu_parity as P

pytestmark = pytest.mark.gpu


def test_quirks_near_collide(ctx):
    P.test_ref_golden_near_collide(ctx, "quirks")
    g = P._ref_golden("ref_near_collide", "quirks")
    tp = g["quirks_obj"] == 8
    assert tp.sum() > 15 and len(np.unique(g["quirks_color"][tp].round(6), axis=0)) > 10   # the texture was looked up


def test_quirks_view_trace(ctx):
    P.test_ref_golden_view_trace(ctx, "quirks")
    g = P._ref_golden("ref_view_trace", "quirks")
    assert g["quirks_counts"].sum() < len(g["quirks_counts"])   # Q8: pixels ending on a non-IsDiff object have no hit point


def test_quirks_photon_pass(ctx):
    P.test_ref_golden_photon_pass(ctx, "quirks")
