#!/usr/bin/env python
"""How would other work-unit shapes change the filtered gather's work?  (no GPU, no kernel change)

    python tools/unit_shape_study.py [--film 1024x768] [--photons 2000000]

The shipped unit is <= 64 consecutive hit points of one run of x-adjacent occupied cells in ONE row of
the hash grid; its deposits are those of 3 x 3 rows x (n + 2) cells.  This script takes the hit points
and deposits of the bench workload from the product's kernels run on the CPU (tests/device_host) and
counts, for units that merge RY x RZ neighbouring rows, the two quantities the gather's time follows
(t = 0.70 us x slices + 0.53 ns x pair slots): deposit visits (/32 = slices) and
visits x hit points of the unit (an upper bound of the pair slots: the bounding-box cull comes after).
Shape 1x1 is the shipped one and calibrates the bound against tools/gather_work_model.py.
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def study(hp_cell, dep_cell, dim, ry, rz, unit_hp=64):
    """hp_cell / dep_cell: integer cell coordinates [n, 3] (x, y, z).  Returns (units, visits, visits x H)."""
    dx, dy, dz = dim
    cnt = np.zeros((dz, dy, dx), dtype=np.int64)
    np.add.at(cnt, (dep_cell[:, 2], dep_cell[:, 1], dep_cell[:, 0]), 1)
    # deposits of the neighbourhood rows of a row group, as prefix sums along x
    gy, gz = (dy + ry - 1) // ry, (dz + rz - 1) // rz
    nb = np.zeros((gz, gy, dx + 1), dtype=np.int64)
    for jz in range(gz):
        z0, z1 = max(jz * rz - 1, 0), min(jz * rz + rz + 1, dz)
        for jy in range(gy):
            y0, y1 = max(jy * ry - 1, 0), min(jy * ry + ry + 1, dy)
            nb[jz, jy, 1:] = np.cumsum(cnt[z0:z1, y0:y1].sum(axis=(0, 1)))
    g = (hp_cell[:, 2] // rz) * gy + hp_cell[:, 1] // ry
    order = np.lexsort((hp_cell[:, 0], g))
    g, x = g[order], hp_cell[order, 0]
    units = visits = pairs = 0
    i, n = 0, len(g)
    while i < n:
        j = i
        # one run: same row group, x-adjacent occupied cells
        while j + 1 < n and g[j + 1] == g[i] and x[j + 1] - x[j] <= 1:
            j += 1
        for b in range(i, j + 1, unit_hp):
            e = min(b + unit_hp, j + 1)
            x0, x1 = max(x[b] - 1, 0), min(x[e - 1] + 1, dx - 1)
            row = nb[g[i] // gy, g[i] % gy]
            v = int(row[x1 + 1] - row[x0])
            units += 1
            visits += v
            pairs += v * (e - b)
        i = j + 1
    return units, visits, pairs


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--film", default="1024x768")
    ap.add_argument("--photons", type=int, default=2000000)
    a = ap.parse_args()
    from device_host import DeviceHost, make_grid
    from raytrace3_b200 import abi
    from raytrace3_b200.scene import make_scene, reference_camera
    w, h = (int(v) for v in a.film.split("x"))
    dh = DeviceHost()
    hs = dh.scene(make_scene("reference", textures="synthetic"))
    prm = abi.default_params()
    prm.newton_mode = abi.SPPM_NEWTON_FAST
    eye, _ = hs.eye_pass_kernel(prm, reference_camera(w, h), pass_id=0)
    dep, _, _ = hs.photon_trace_wavefront(prm, 0, a.photons, 1, step_blocks=64, newton_blocks=64)
    pos = np.ascontiguousarray(eye["pos"])
    g = make_grid(pos, prm.init_r2)
    org, dim = np.array(g.org[:]), np.array(g.dim[:])
    cell = lambda p: np.clip(np.floor((p - org) * g.inv_cell).astype(np.int64), 0, dim - 1)
    x = dep["x"]
    keep = ((x >= np.array(g.keep_lo[:])) & (x <= np.array(g.keep_hi[:]))).all(axis=1)
    hc, dc = cell(pos), cell(x[keep])
    base = None
    for ry, rz in ((1, 1), (2, 1), (1, 2), (2, 2), (3, 3)):
        u, v, p = study(hc, dc, dim, ry, rz)
        base = base or (v, p)
        print(json.dumps(dict(rows=f"{ry}x{rz}", ranges_per_unit=(ry + 2) * (rz + 2), units=u, hp_per_unit=round(len(hc) / u, 1),
                              visits=v, slices=v // 32, visits_x_hp=p, visits_vs_shipped=round(v / base[0], 3),
                              pairs_vs_shipped=round(p / base[1], 3))))


if __name__ == "__main__":
    main()
