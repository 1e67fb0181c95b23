"""Regenerate data files derived from the reference checkout (run in the dev container only:
/root/reference does not exist on the GPU box).

  raytrace3_b200/data/teapot32.npy — the 32 bicubic patches of raytracer/teapot.bpt (Newell's
  Utah-teapot data set), float64 [32][16][3], raw file coordinates, parsed exactly as the
  Bezier3Obj constructor does (Bezier.h:204-218).
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from raytrace3_b200.scene import DATA_DIR, read_bpt  # noqa: E402

if __name__ == "__main__":
    src = "/root/reference/raytracer/teapot.bpt"
    P = read_bpt(src)
    assert P.shape == (32, 16, 3)
    os.makedirs(DATA_DIR, exist_ok=True)
    np.save(os.path.join(DATA_DIR, "teapot32.npy"), P)
    print("teapot32.npy", P.shape, P.min(axis=(0, 1)), P.max(axis=(0, 1)))
