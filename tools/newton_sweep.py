"""Parity sweep of the Newton arithmetic modes on the GPU: FAST (monomial Horner + FMA) and
FAST_BERNSTEIN against EXACT (which is bit-identical to the oracle / the reference) on many
(ray, patch) pairs with the counter-RNG seeds.  Prints hit/miss disagreements and the error
distribution of (t, u, v) — the north-star bar is 1e-6 relative."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from conftest import rays_at_patches, rays_at_teapot  # noqa: E402
from raytrace3_b200 import Context, abi  # noqa: E402
from raytrace3_b200.scene import make_scene  # noqa: E402

if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000
    sc = make_scene("reference", textures=None)
    patch = (np.arange(n) % 32).astype(np.int32)
    rays = rays_at_patches(patch, 77, sc)
    rays[: n // 4] = rays_at_teapot(n // 4, 78, sc)
    ctx = Context(0)
    ctx.upload_scene(sc)
    res = {}
    for name, mode in (("exact", abi.SPPM_NEWTON_EXACT), ("fast", abi.SPPM_NEWTON_FAST), ("fast_bernstein", abi.SPPM_NEWTON_FAST_BERNSTEIN)):
        p = abi.default_params()
        p.newton_mode = mode
        ctx.set_params(p)
        t0 = time.perf_counter()
        res[name] = ctx.bezier_patch_batch(rays, patch, None)
        print(f"{name:15s} {time.perf_counter() - t0:.2f} s (incl. host copies)")
    ref = res["exact"]
    hr = ref[:, 0] < 1e19
    print("pairs", n, "solutions", int(hr.sum()))
    for name in ("fast", "fast_bernstein"):
        got = res[name]
        hg = got[:, 0] < 1e19
        both = hr & hg
        rel = np.abs(got[both] - ref[both]) / np.maximum(np.abs(ref[both]), 1e-12)
        ab = np.abs(got[both] - ref[both])
        print(f"{name:15s} hit/miss disagreements {int((hr != hg).sum())}; rel err max {rel.max():.3e} p99.99 {np.quantile(rel, 0.9999):.3e} "
              f"median {np.median(rel):.3e}; abs err max {ab.max():.3e}; pairs over 1e-6 rel: {int((rel.max(axis=1) > 1e-6).sum())}")
