"""Generate tests/golden/*.npz — golden input/output vectors of the CPU oracle (the reference
ships none).  Deterministic; rerun only when the oracle's semantics change."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import rays_at_patches, rays_at_teapot, rays_in_room  # noqa: E402
from oracle import Oracle  # noqa: E402
from raytrace3_b200 import abi  # noqa: E402
from raytrace3_b200.scene import make_scene, reference_camera  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")

if __name__ == "__main__":
    os.makedirs(GOLD, exist_ok=True)
    o = Oracle()
    prm = abi.default_params()
    ref = make_scene("reference", textures=None)
    cor = make_scene("cornell", textures=None)
    rng = np.random.RandomState(123)
    patch = rng.randint(0, 32, 96).astype(np.int32)
    rays = rays_at_patches(patch, 21, ref)
    rays[:16] = rays_at_teapot(16, 21, ref)
    seeds = rng.uniform(size=(96, 50, 2))
    np.savez_compressed(os.path.join(GOLD, "newton.npz"), rays=rays, patch=patch, seeds=seeds,
                        tuv=o.bezier_patch_batch(ref, prm, rays, patch, seeds))
    rays = np.concatenate([rays_in_room(300, 31), rays_at_teapot(212, 32, ref)])
    ids = np.arange(512, dtype=np.uint64) * 5 + 1
    h = o.intersect_batch(ref, prm, rays, ids, pass_id=2, node=5)
    np.savez_compressed(os.path.join(GOLD, "near_collide.npz"), rays=rays, ids=ids, **{k: h[k] for k in h.dtype.names})
    hp = o.eye_pass(cor, prm, reference_camera(32, 24), 0)
    np.savez_compressed(os.path.join(GOLD, "eye_cornell.npz"), **hp)
    r = o.photon_pass(cor, prm, 0, 500, 3, hp=hp, keep_deposits=True)
    np.savez_compressed(os.path.join(GOLD, "photon_cornell.npz"), newN=r["newN"], tao=r["tao"], dep_x=r["dep"]["x"],
                        dep_n=r["dep"]["n"], dep_flux=r["dep"]["flux"], dep_photon=r["dep"]["photon"],
                        dep_bounce=r["dep"]["bounce"], **{"hp_" + k: v for k, v in hp.items()})
    for f in sorted(os.listdir(GOLD)):
        print(f, os.path.getsize(os.path.join(GOLD, f)))
