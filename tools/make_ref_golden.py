"""Generate tests/golden/ref_*.npz — golden input/output vectors produced by THE REFERENCE'S OWN
CODE (oracle/_ref/libref.so = /root/reference/raytracer/*.h compiled against the OpenCV/FLANN
shim, oracle/ref_shim/build_ref.py).  The reference ships no tests or fixtures and
neither /root/reference nor, in a fresh clone, the git-ignored libref.so exist where the test
suites run, so its outputs are frozen here: tests/test_ref_golden_cpu.py checks the oracle
against them bit for bit and tests/test_gpu_parity.py checks the CUDA path against them.

Run in the dev container only (needs /root/reference):   python tools/make_ref_golden.py

Random numbers: the reference's generator is not a parity target, so its draws are injected.
For scene-level functions the oracle first runs the same input with its counter RNG and logs
the draws (orc_log_begin/end); the reference replays exactly those values in the order it
consumes them.  A stored vector is therefore a function of (inputs, seed, ids) alone and any
implementation using the same counter RNG must reproduce it.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import custom_scene, quirks_scene, rays_at_patches, rays_at_teapot, rays_in_room  # noqa: E402
from oracle import Oracle  # noqa: E402
from oracle import ref as ref_mod  # noqa: E402
from raytrace3_b200 import abi  # noqa: E402
from raytrace3_b200.scene import make_scene, reference_camera, teapot_patches  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
HIT_FIELDS = ("obj", "inside", "t", "pos", "n", "color")
SCENES = {
    "cornell": lambda: make_scene("cornell", textures=None),
    "reference": lambda: make_scene("reference", textures="synthetic"),
    "stress": lambda: make_scene("stress", textures=None),
    "custom": custom_scene,
    "quirks": quirks_scene,   # textured teapot (Q6), materials with a zero diffuse channel (Q8): stored in ref_quirks_*.npz
}


def pieces(ref):
    """Bezier evaluation, AABB test, sampling, optics, camera, texture, update, tone map, resolve."""
    rng = np.random.RandomState(100)
    P = teapot_patches()
    out = {}
    n = 240
    idx, scale = rng.randint(0, 32, n), rng.uniform(0.5, 5, n)
    uv = rng.uniform(-0.2, 1.2, (n, 2))
    ev = np.array([np.stack(ref.bezier_eval(P[idx[k]] * scale[k], uv[k, 0], uv[k, 1])) for k in range(n)])
    out.update(be_idx=idx, be_scale=scale, be_uv=uv, be_out=ev)
    n = 1500
    box, org, dirs = np.zeros((n, 6)), np.zeros((n, 3)), np.zeros((n, 3))
    for k in range(n):
        lo = rng.uniform(-5, 5, 3)
        box[k] = np.concatenate([lo, lo + rng.uniform(0.1, 6, 3)])
        org[k] = rng.uniform(-12, 12, 3)
        d = rng.normal(size=3)
        if k % 2 == 0:
            d = box[k, :3] + rng.uniform(size=3) * (box[k, 3:] - box[k, :3]) - org[k]  # aimed into the box
        if k % 5 == 0:
            d[rng.randint(3)] = rng.choice([0.0, 5e-5, -5e-5, 1e-4])  # IsZero(dir) edge
        if k % 7 == 0:
            org[k] = lo + rng.uniform(0, 0.1, 3)                       # origin inside / on the box (Q5)
        dirs[k] = d / np.linalg.norm(d)
    out.update(ab_box=box, ab_org=org, ab_dir=dirs,
               ab_hit=np.array([ref.aabb_intersect(box[k], org[k], dirs[k]) for k in range(n)]))
    n = 300
    nn, u12, refl = np.zeros((n, 3)), np.zeros((n, 2)), np.zeros((n, 3))
    I, N, rf = np.zeros((n, 3)), np.zeros((n, 3)), np.zeros((n, 2, 3))
    for k in range(n):
        v = rng.normal(size=3)
        v /= np.linalg.norm(v)
        if k % 10 == 0:
            v = np.array([0.0, 0.0, rng.choice([-1.0, 1.0])])          # GetAnormal special case (Vec3.h:87)
        u12[k] = rng.uniform(size=2)
        if k % 17 == 0:
            u12[k, 0] = 1.0 - 1e-9                                     # theta below 1e-4: rotated() returns *this
        nn[k] = v
        refl[k] = ref.unit_rand_refl(v, u12[k, 0], u12[k, 1])
        i = rng.normal(size=3)
        i /= np.linalg.norm(i)
        I[k], N[k] = i, (v if i @ v < 0 else -v)
        rf[k, 0], rf[k, 1] = ref.refracted(I[k], N[k], 1.5), ref.refracted(I[k], N[k], 1 / 1.5)
    out.update(sr_n=nn, sr_u=u12, sr_out=refl, rf_I=I, rf_N=N, rf_out=rf)
    cams = []
    for w, h in [(320, 240), (1024, 768), (64, 48)]:
        pos = np.array([50., 35., 230.]) + rng.uniform(-1, 1, 3) * (w == 64)
        look = pos + np.array([0, 0.042612, -1.])
        rc, ray = ref.camera(pos, look, w, h, y=h // 3, x=w // 5)
        cams.append(np.concatenate([pos, look, [w, h, h // 3, w // 5], list(rc.pos), list(rc.dir), list(rc.du), list(rc.dv), ray]))
    out["cam"] = np.array(cams)
    tex = rng.randint(0, 256, size=(13, 7, 3)).astype(np.uint8)
    n = 300
    tuv = rng.uniform(-3, 3, (n, 2))
    for k in range(0, n, 9):
        tuv[k] = (float(rng.randint(-2, 3)), rng.choice([0.0, -1e-20, 1 - 1e-16]))
    out.update(tx_map=tex, tx_uv=tuv, tx_out=np.array([ref.color_uv(tex, *tuv[k]) for k in range(n)]))
    n = 400
    newN, Nn = rng.randint(0, 40, n).astype(np.int32), rng.randint(0, 30, n).astype(np.int32)
    Nn[:100] = 0                                                       # the as-shipped case (Q1)
    r2, tao = rng.uniform(0.1, 2, n), rng.uniform(0, 50, (n, 3))
    b = [x.copy() for x in (newN, Nn, r2, tao)]
    ref.update(*b)
    out.update(up_newN=newN, up_N=Nn, up_r2=r2, up_tao=tao, up_o_newN=b[0], up_o_N=b[1], up_o_r2=b[2], up_o_tao=b[3])
    x = np.concatenate([rng.uniform(0, 3, 200), [0.0, 1e-12, 50.0, 0.002, 0.0019]])
    out.update(ti_x=x, ti_out=np.array([ref.to_int(v) for v in x], np.int32))
    rs = np.concatenate([rng.uniform(0, 60, (40, 3)), rng.uniform(0.05, 2, (40, 1))], axis=1)
    out.update(rs_in=rs, rs_out=np.array([ref.resolve_one(r[:3], r[3], 7, 100) for r in rs]))
    return out


def newton(ref):
    sc = make_scene("reference", textures=None)
    rng = np.random.RandomState(101)
    n = 96
    patch = rng.randint(0, 32, n).astype(np.int32)
    rays = rays_at_patches(patch, 41, sc)
    rays[:20] = rays_at_teapot(20, 42, sc)
    seeds = rng.uniform(size=(n, 50, 2))
    P = sc.objvec[8].P
    tuv = np.array([ref.newton(P[patch[i]], rays[i, :3], rays[i, 3:], seeds[i]) for i in range(n)])
    # flat patch: singular Jacobian (inv() -> zeros), a negative-t solution, rays in the plane
    g = np.zeros((4, 4, 3))
    for i in range(4):
        for k in range(4):
            g[i, k] = (k / 3.0, i / 3.0, 0.0)
    fr = np.array([[0.25, 0.75, 2.0, 0, 0, -1.0], [0.25, 0.75, -2.0, 0, 0, -1.0], [0.25, 0.75, 2.0, 1.0, 0, 0],
                   [0.5, 0.5, 0.005, 1.0, 0, 0]])
    fs = np.full((50, 2), 0.5)
    ft = np.array([ref.newton(g.reshape(16, 3), r[:3], r[3:], fs) for r in fr])
    return dict(rays=rays, patch=patch, seeds=seeds, tuv=tuv, flat_P=g.reshape(16, 3), flat_rays=fr, flat_tuv=ft)


def near_collide(orc, ref, prm, names=("cornell", "reference", "stress", "custom")):
    out = {}
    for name in names:
        sc = SCENES[name]()
        h = ref.scene(sc)
        n_room, n_tp = (160, 0) if name in ("cornell", "custom") else (80, 100)
        rays = rays_in_room(n_room, 43)
        if n_tp:
            rays = np.concatenate([rays, rays_at_teapot(n_tp, 44, sc)])
        rec = {f: [] for f in HIT_FIELDS}
        for i in range(rays.shape[0]):
            orc.log_begin()
            orc.intersect_batch(sc, prm, rays[i:i + 1], np.array([i], np.uint64), pass_id=1, node=2)
            vals, kinds = orc.log_end()
            b = ref.near_collide(h, rays[i], vals, kinds)
            for f in HIT_FIELDS:
                rec[f].append(np.array(b[f]))
        out[name + "_rays"] = rays
        for f in HIT_FIELDS:
            out[f"{name}_{f}"] = np.array(rec[f])
    return out


def photon(orc, ref, prm, names=("cornell", "reference", "custom")):
    """Light::emit + PhotonTrace + radiusSearch/accumulate, one photon per light per call, on the
    hit points of a 64x48 eye pass (stored: they are the input).  Per photon the totals, and the
    accumulation over all photons in photon order (numpy float64 adds, so a checker can repeat it)."""
    out = {}
    for name in names:
        sc = SCENES[name]()
        h = ref.scene(sc)
        hp = orc.eye_pass(sc, prm, reference_camera(64, 48), 0)
        n = hp["pos"].shape[0]
        acc_n, acc_t = np.zeros(n, np.int64), np.zeros((n, 3))
        tot_n, tot_t = [], []
        for j in range(300):
            orc.log_begin()
            orc.photon_pass(sc, prm, j, 1, 5, hp=hp, use_kdtree=False)
            vals, kinds = orc.log_end()
            newN, tao = ref.photon(h, len(sc.plvec), hp, vals, kinds)
            acc_n += newN
            acc_t += tao
            tot_n.append(int(newN.sum()))
            tot_t.append(tao.sum(axis=0))
        assert acc_n.sum() > 50, (name, acc_n.sum())
        out.update({f"{name}_hp_{k}": hp[k] for k in ("pos", "n", "wgt", "r2", "pix")})
        out.update({f"{name}_acc_newN": acc_n, f"{name}_acc_tao": acc_t, f"{name}_tot_newN": np.array(tot_n),
                    f"{name}_tot_tao": np.array(tot_t)})
    return out


def view_trace(orc, ref, prm, names=("reference", "custom")):
    out = {}
    w, hh = 40, 30
    for name in names:
        sc = SCENES[name]()
        h = ref.scene(sc)
        cam = reference_camera(w, hh)
        rec = {f: [] for f in ("pos", "n", "wgt", "r2", "pix")}
        counts, pixels = [], []
        for i in range(0, hh, 2):
            for j in range(0, w, 3):
                orc.log_begin()
                orc.view_trace_pixel(sc, prm, cam, i, j)
                vals, kinds = orc.log_end()
                b = ref.view_trace(h, cam.pos, cam.look, w, hh, i, j, vals, kinds)
                for f in rec:
                    rec[f].append(b[f])
                counts.append(b["pos"].shape[0])
                pixels.append((i, j))
        out[name + "_pixels"] = np.array(pixels, np.int32)
        out[name + "_counts"] = np.array(counts, np.int32)
        for f in rec:
            out[f"{name}_{f}"] = np.concatenate(rec[f])
    out["film"] = np.array([w, hh], np.int32)
    return out


if __name__ == "__main__":
    if not ref_mod.available():
        sys.exit("oracle/_ref/libref.so is missing: run oracle/ref_shim/build_ref.py where /root/reference exists")
    os.makedirs(GOLD, exist_ok=True)
    orc, ref, prm = Oracle(), ref_mod.Reference(), abi.default_params()
    only = sys.argv[1:]          # e.g. `make_ref_golden.py ref_quirks_photon` regenerates one file
    q = ("quirks",)   # same keys as the three files above, so in files of their own
    for name, make in (("ref_pieces", lambda: pieces(ref)), ("ref_newton", lambda: newton(ref)),
                       ("ref_near_collide", lambda: near_collide(orc, ref, prm)), ("ref_photon", lambda: photon(orc, ref, prm)),
                       ("ref_view_trace", lambda: view_trace(orc, ref, prm)),
                       ("ref_quirks_near_collide", lambda: near_collide(orc, ref, prm, q)),
                       ("ref_quirks_photon", lambda: photon(orc, ref, prm, q)),
                       ("ref_quirks_view_trace", lambda: view_trace(orc, ref, prm, q))):
        if only and name not in only:
            continue
        data = make()
        path = os.path.join(GOLD, name + ".npz")
        np.savez_compressed(path, **data)
        print(name, os.path.getsize(path), "bytes,", len(data), "arrays")
