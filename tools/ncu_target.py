#!/usr/bin/env python
"""One region of the hot path between cudaProfilerStart/Stop, for `ncu --profile-from-start off`.

  ncu --set full --clock-control none --import-source on --profile-from-start off \
      -k regex:newton_items -f -o gpurun_out/x python tools/ncu_target.py --region round

Regions: build (sppm_build_index after an eye pass), round (one PhotonMap round after warm-up rounds:
step / Newton / binning / gather / update kernels), patch (the M2 ray-patch microbenchmark), eye.
The set-up and warm-up run outside the profiled range, so -s/-c offsets do not depend on them.
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402

import bench  # noqa: E402
from raytrace3_b200 import Context, abi  # noqa: E402
from raytrace3_b200.scene import make_scene, reference_camera  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--region", default="round", choices=["build", "round", "patch", "eye"])
    ap.add_argument("--workload", default=bench.DEFAULT_WORKLOAD)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--rounds", type=int, default=1)
    ap.add_argument("--newton", default="fast")
    ap.add_argument("--update", default="reference")
    ap.add_argument("--photons", type=int, default=0)
    args = ap.parse_args()
    wl = bench.WORKLOADS[args.workload]
    prm = bench.make_params(abi, fast=args.newton == "fast")
    prm.count_pairs = 0
    if args.update == "progressive":
        prm.update_mode = abi.SPPM_UPDATE_PROGRESSIVE
    ctx = Context(0, prm)
    scene = make_scene(wl["scene"])
    cam = reference_camera(wl["w"], wl["h"])
    onetime = args.photons or wl["onetime"]
    ctx.upload_scene(scene)
    prof = torch.cuda.profiler
    if args.region == "eye":
        ctx.eye_pass(cam, 0)
        ctx.synchronize()
        prof.start()
        ctx.eye_pass(cam, 1)
        ctx.synchronize()
        prof.stop()
        return
    ctx.eye_pass(cam, 0)
    ctx.build_index()
    ctx.synchronize()
    if args.region == "build":
        ctx.eye_pass(cam, 1)
        ctx.synchronize()
        prof.start()
        ctx.build_index()
        ctx.synchronize()
        prof.stop()
        return
    if args.region == "patch":
        ctx.bezier_patch_bench(1 << 18, 1)
        ctx.synchronize()
        prof.start()
        ctx.bezier_patch_bench(1 << 18, 1)
        ctx.synchronize()
        prof.stop()
        return
    for i in range(args.warmup):
        ctx.photon_pass(0, onetime, 1 + i)
        ctx.update()
    ctx.synchronize()
    prof.start()
    for i in range(args.rounds):
        ctx.photon_pass(0, onetime, 1 + args.warmup + i)
        ctx.update()
        ctx.synchronize()
    prof.stop()
    print("counters", ctx.counters(), file=sys.stderr)


if __name__ == "__main__":
    main()
