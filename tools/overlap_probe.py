"""Probe: how much aggregate throughput does one B200 give when two independent contexts
(two streams) run photon rounds concurrently?  >1x means kernels from different streams fill
each other's idle issue slots, i.e. pipelining trace(i+1) against gather(i) would pay."""
import os
import sys
import threading
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from raytrace3_b200 import Context, abi  # noqa: E402
from raytrace3_b200.scene import make_scene, reference_camera  # noqa: E402


def setup():
    prm = abi.default_params()
    prm.newton_mode = abi.SPPM_NEWTON_FAST
    ctx = Context(0, prm)
    ctx.upload_scene(make_scene("reference"))
    ctx.eye_pass(reference_camera(1024, 768), 0)
    ctx.build_index()
    return ctx


def rounds(ctx, n, base):
    for i in range(n):
        ctx.photon_pass(0, 2_000_000, base + i)
        ctx.update()
    ctx.synchronize()


if __name__ == "__main__":
    a, b = setup(), setup()
    rounds(a, 3, 0)
    rounds(b, 3, 100)
    n = 30
    t0 = time.perf_counter()
    rounds(a, n, 10)
    t1 = time.perf_counter() - t0
    print(f"one context : {n * 2e6 / t1 / 1e6:.1f} M photons/s ({1e3 * t1 / n:.2f} ms/round)")
    th = [threading.Thread(target=rounds, args=(c, n, 200 + 100 * k)) for k, c in enumerate((a, b))]
    t0 = time.perf_counter()
    for t in th:
        t.start()
    for t in th:
        t.join()
    t2 = time.perf_counter() - t0
    print(f"two contexts: {2 * n * 2e6 / t2 / 1e6:.1f} M photons/s aggregate ({1e3 * t2 / n:.2f} ms per pair of rounds) -> x{(2 * n / t2) / (n / t1):.2f}")
