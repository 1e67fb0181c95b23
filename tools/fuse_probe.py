"""Times RayTracer.PhotonMap on the reference's as-shipped loop shape (Raytracer.h:218, 384:
rounds of onetime = 100 photons) round by round and with the rounds fused into one photon pass.
Usage (GPU box): python tools/fuse_probe.py [rounds_unfused] [rounds_fused]"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from raytrace3_b200 import Context, RayTracer, abi  # noqa: E402
from raytrace3_b200.scene import make_scene, reference_camera  # noqa: E402

n_unfused = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
n_fused = int(sys.argv[2]) if len(sys.argv) > 2 else 10000
out = {}
with Context(0, abi.default_params()) as ctx:
    sc = make_scene("reference", textures=None)
    cam = reference_camera(320, 240)
    for name, fuse, cnt in (("warm", True, 50), ("round_by_round", False, n_unfused), ("fused", True, n_fused)):
        rt = RayTracer(ctx)
        rt.camera, rt.scene, rt.onetime, rt.fuse_rounds = cam.copy(), sc, 100, fuse
        rt.GetHitPoint()
        ctx.build_index()
        ctx.synchronize()
        t0 = time.perf_counter()
        rt.PhotonMap(cnt)            # ends with the film read-back: synchronous
        dt = time.perf_counter() - t0
        out[name] = {"rounds": cnt, "photons": cnt * 100, "s": dt, "photons_per_s": cnt * 100 / dt, "film_sum": float(rt.camera.pic.sum())}
print(json.dumps(out))
