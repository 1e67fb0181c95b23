"""`python -m raytrace3_b200.render SCENE [options]` — the Python twin of examples/sppm_render
(SURVEY §8f-2): renders a scene description file on the GPU and writes PPM or JPEG (nvJPEG)."""
from __future__ import annotations

import argparse

import numpy as np

from . import Context, abi
from .scene import Texture
from .scenefile import load_scene_file, render_job


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__)
    ap.add_argument("scene")
    ap.add_argument("--size")
    for k in ("passes", "replicas", "rounds", "photons", "quality"):
        ap.add_argument(f"--{k}", type=int)
    ap.add_argument("--seed", type=lambda s: int(s, 0))
    ap.add_argument("--shared", action="store_true")
    ap.add_argument("--newton", choices=["exact", "fast"])
    ap.add_argument("--gather", choices=["exact", "filtered"])
    ap.add_argument("--math", choices=["exact", "fast"])
    ap.add_argument("--device", type=int, default=0)
    ap.add_argument("--out")
    ap.add_argument("--format", choices=["ppm", "jpg"])
    a = ap.parse_args(argv)
    ctx = Context(a.device)

    def tex(path):  # JPEG textures are decoded on the device (nvJPEG); .bgr8 / others by the host class
        with open(path, "rb") as f:
            data = f.read()
        return Texture(ctx.jpeg_decode(data)) if data[:2] == b"\xff\xd8" else Texture(path)

    scene, job = load_scene_file(a.scene, texture_loader=tex)
    if a.size:
        job.w, job.h = (int(x) for x in a.size.lower().split("x"))
    for k in ("passes", "replicas", "rounds", "photons", "quality", "seed"):
        if getattr(a, k) is not None:
            setattr(job, k, getattr(a, k))
    job.shared = job.shared or a.shared
    if a.newton:
        job.newton_mode = abi.SPPM_NEWTON_FAST if a.newton == "fast" else abi.SPPM_NEWTON_EXACT
    if a.gather:
        job.gather_mode = abi.SPPM_GATHER_EXACT if a.gather == "exact" else abi.SPPM_GATHER_FILTERED
    if a.math:
        job.math_mode = abi.SPPM_MATH_FAST if a.math == "fast" else abi.SPPM_MATH_EXACT
    if a.out:
        job.out_prefix = a.out
    if a.format:
        job.format = a.format
    film, img = render_job(ctx, scene, job)
    path = f"{job.out_prefix}.{job.format}"
    if job.format == "jpg":
        with open(path, "wb") as f:
            f.write(ctx.jpeg_encode(img, job.quality))
    else:
        with open(path, "wb") as f:
            f.write(f"P6\n{job.w} {job.h}\n255\n".encode())
            f.write(np.ascontiguousarray(img[:, :, ::-1]).tobytes())  # film channels are BGR (Q11)
    print(f"film sum {film.sum():.17g}")
    print(f"save:{path}")


if __name__ == "__main__":
    main()
