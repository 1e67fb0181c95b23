#!/usr/bin/env python
"""Algorithmic work of the filtered gather on a BASELINE-shaped workload, without a GPU.

    python tools/gather_work_model.py [--film 1024x768] [--photons 2000000] [--scene reference]
    SPPM_DEVICE_HOST_EXTRA="-DSPPM_UNIT_HP=32" python tools/gather_work_model.py ...   # a compile-time knob

Runs the product's own kernels on the CPU (tests/device_host: eye pass -> wavefront photon pass ->
hash grid + work units -> gather_filtered_kernel) and prints the counters that do not depend on the
hardware: how often a deposit is streamed to a work unit, what the bounding-box cull keeps, how many
(hit point, deposit) pairs reach the FP32 classification per candidate pair (= pair inside the search
ball) and per accepted pair, how many pair slots are issued for them (padding, idle lanes) and how
often a lane falls back to the careful path.  The classification loop costs about 15 warp
instructions per 32 pair slots x ..., so `pair_slots` is the number the loop's
time follows; use this to compare unit shapes / cull rules before spending GPU minutes on them.
Prints one JSON line.
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--film", default="1024x768")
    ap.add_argument("--photons", type=int, default=2000000)
    ap.add_argument("--scene", default="reference", choices=["reference", "cornell", "stress"])
    ap.add_argument("--blocks", type=int, default=64)
    a = ap.parse_args()
    from device_host import DeviceHost
    from raytrace3_b200 import abi
    from raytrace3_b200.scene import make_scene, reference_camera
    w, h = (int(v) for v in a.film.split("x"))
    dh = DeviceHost()
    sc = make_scene(a.scene, textures="synthetic" if a.scene != "cornell" else None)
    hs = dh.scene(sc)
    prm = abi.default_params()
    prm.newton_mode = abi.SPPM_NEWTON_FAST
    t0 = time.time()
    eye, _ = hs.eye_pass_kernel(prm, reference_camera(w, h), pass_id=0)
    hp = dict(pos=eye["pos"], n=eye["n"], wgt=eye["wgt"], r2=eye["r2"])
    t1 = time.time()
    dep, cnt, _ = hs.photon_trace_wavefront(prm, 0, a.photons, 1, step_blocks=a.blocks, newton_blocks=a.blocks)
    t2 = time.time()
    dh.gather_work()
    newN, tao, info, pairs = dh.gather_filtered(hp, dep, prm.init_r2, blocks=a.blocks)
    t3 = time.time()
    wk = dh.gather_work()
    out = dict(workload=f"{a.scene}-{w}x{h}", photons=a.photons, extra=os.environ.get("SPPM_DEVICE_HOST_EXTRA", ""),
               hit_points=int(eye.shape[0]), deposits=int(dep["x"].shape[0]), deposits_kept=info["kept"], units=info["n_units"],
               work_items=info["n_items"], cand_pairs=pairs[0], acc_pairs=pairs[1], newN_sum=int(newN.sum()), **wk)
    out["visits_per_kept_deposit"] = round(wk["visits"] / max(info["kept"], 1), 3)
    out["cull_keeps"] = round(wk["kept"] / max(wk["visits"], 1), 4)
    out["pairs_per_cand"] = round(wk["pairs"] / max(pairs[0], 1), 3)
    out["pair_slots_per_cand"] = round(wk["pair_slots"] / max(pairs[0], 1), 3)
    out["redo_lanes_per_slice"] = round(wk["redo_lanes"] / max(wk["slices"], 1), 5)
    out["hp_per_unit"] = round(eye.shape[0] / max(info["n_units"], 1), 2)
    out["seconds"] = dict(eye=round(t1 - t0, 1), photons=round(t2 - t1, 1), grid_and_gather=round(t3 - t2, 1))
    print(json.dumps(out))


if __name__ == "__main__":
    main()
