"""world_size-2 gloo test of the multi-GPU scheme on CPU: each rank traces a
disjoint photon range over a replicated hit-point map, the per-round accumulators are summed
with all_reduce, and the result equals the single-rank run (newN bit-exact, tao to rounding).
The compute here is the CPU oracle — the test covers the host-side sharding logic."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, onetime, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ["OMP_NUM_THREADS"] = "1"
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import Oracle
    from raytrace3_b200 import abi, photon_range
    from raytrace3_b200.scene import make_scene, reference_camera
    orc = Oracle()
    prm = abi.default_params()
    sc = make_scene("cornell", textures=None)
    hp = orc.eye_pass(sc, prm, reference_camera(32, 24), 0)   # replicated: same on every rank
    n = hp["pos"].shape[0]
    newN, tao = np.zeros(n, np.int32), np.zeros((n, 3))
    for rnd in range(2):
        b, c = photon_range(onetime, rank, world)
        r = orc.photon_pass(sc, prm, b, c, rnd, hp=hp)
        dN, dtao = torch.from_numpy(r["newN"].copy()), torch.from_numpy(r["tao"].copy())
        dist.all_reduce(dN)      # int32: exact
        dist.all_reduce(dtao)    # float64: order-dependent rounding only
        newN += dN.numpy()
        tao += dtao.numpy()
    if rank == 0:
        q.put((newN, tao))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_photon_pass_equals_single_rank():
    onetime, world = 1501, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, onetime, q)) for r in range(world)]
    for p in procs:
        p.start()
    newN, tao = q.get(timeout=300)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    sys.path.insert(0, ROOT)
    from oracle import Oracle
    from raytrace3_b200 import abi
    from raytrace3_b200.scene import make_scene, reference_camera
    orc = Oracle()
    prm = abi.default_params()
    sc = make_scene("cornell", textures=None)
    hp = orc.eye_pass(sc, prm, reference_camera(32, 24), 0)
    n = hp["pos"].shape[0]
    ref_N, ref_tao = np.zeros(n, np.int32), np.zeros((n, 3))
    for rnd in range(2):
        r = orc.photon_pass(sc, prm, 0, onetime, rnd, hp=hp)
        ref_N += r["newN"]
        ref_tao += r["tao"]
    assert ref_N.sum() > 0
    np.testing.assert_array_equal(newN, ref_N)
    np.testing.assert_allclose(tao, ref_tao, rtol=1e-12, atol=0)


def _fused_worker(rank, world, port, cnt, onetime, q):
    """RayTracer.PhotonMap's sharded, fused form: every rank traces its share of the fused id range
    under the eye pass's id, ONE all_reduce per call."""
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ["OMP_NUM_THREADS"] = "1"
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import Oracle
    from raytrace3_b200 import abi, photon_calls
    from raytrace3_b200.scene import make_scene, reference_camera
    orc = Oracle()
    prm = abi.default_params()
    sc = make_scene("cornell", textures=None)
    hp = orc.eye_pass(sc, prm, reference_camera(32, 24), 0)
    n = hp["pos"].shape[0]
    newN, tao = np.zeros(n, np.int32), np.zeros((n, 3))
    calls = photon_calls(cnt, onetime, True, rank, world)
    for b, c in calls:
        r = orc.photon_pass(sc, prm, b, c, 5, hp=hp)
        dN, dtao = torch.from_numpy(r["newN"].copy()), torch.from_numpy(r["tao"].copy())
        dist.all_reduce(dN)
        dist.all_reduce(dtao)
        newN += dN.numpy()
        tao += dtao.numpy()
    if rank == 0:
        q.put((newN, tao, len(calls)))
    dist.barrier()
    dist.destroy_process_group()


def test_fused_sharded_photon_map_equals_round_by_round():
    """The rounds of an eye pass fused into one call and split over 2 ranks trace the same photons as
    one rank going round by round (photon k of round i = id i*onetime + k under the eye pass's id)."""
    cnt, onetime, world = 5, 301, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_fused_worker, args=(r, world, port, cnt, onetime, q)) for r in range(world)]
    for p in procs:
        p.start()
    newN, tao, n_calls = q.get(timeout=300)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert n_calls == 1
    sys.path.insert(0, ROOT)
    from oracle import Oracle
    from raytrace3_b200 import abi, photon_calls
    from raytrace3_b200.scene import make_scene, reference_camera
    orc = Oracle()
    prm = abi.default_params()
    sc = make_scene("cornell", textures=None)
    hp = orc.eye_pass(sc, prm, reference_camera(32, 24), 0)
    n = hp["pos"].shape[0]
    ref_N, ref_tao = np.zeros(n, np.int32), np.zeros((n, 3))
    rounds = photon_calls(cnt, onetime, False)
    assert rounds == [(i * onetime, onetime) for i in range(cnt)]
    for b, c in rounds:
        r = orc.photon_pass(sc, prm, b, c, 5, hp=hp)
        ref_N += r["newN"]
        ref_tao += r["tao"]
    assert ref_N.sum() > 0
    np.testing.assert_array_equal(newN, ref_N)
    np.testing.assert_allclose(tao, ref_tao, rtol=1e-12, atol=0)


def _eye_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ["OMP_NUM_THREADS"] = "1"
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import Oracle
    from raytrace3_b200 import abi, row_range
    from raytrace3_b200.scene import make_scene, reference_camera
    orc = Oracle()
    prm = abi.default_params()
    sc = make_scene("cornell", textures=None)
    cam = reference_camera(24, 18)
    r0, r1 = row_range(cam.h, rank, world)
    mine = [orc.view_trace_pixel(sc, prm, cam, i, j, 0) for i in range(r0, r1) for j in range(cam.w)]
    share = {k: np.concatenate([m[k] for m in mine]) for k in ("pos", "n", "wgt", "r2", "pix")}
    parts = [None] * world
    dist.all_gather_object(parts, share)          # the exchange sppm_allgather_hitpoints does over NCCL
    if rank == 0:
        q.put(({k: np.concatenate([p[k] for p in parts]) for k in share}, (r0, r1)))
    dist.barrier()
    dist.destroy_process_group()


def test_row_sharded_eye_pass_equals_single_rank():
    """SURVEY §8f-4 on CPU: every rank traces its film rows (draws are addressed by pixel), the
    shares are exchanged, and the union is the hit-point map of the whole eye pass."""
    sys.path.insert(0, ROOT)
    from raytrace3_b200 import row_range
    assert [row_range(10, r, 3) for r in range(3)] == [(0, 3), (3, 6), (6, 10)]
    assert row_range(5, 7, 8) == (4, 5) and row_range(5, 0, 8) == (0, 0)
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_eye_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got, rows0 = q.get(timeout=300)
    for p in procs:
        p.join(timeout=300)
        assert p.exitcode == 0
    assert rows0 == (0, 9)
    from oracle import Oracle
    from raytrace3_b200 import abi
    from raytrace3_b200.scene import make_scene, reference_camera
    whole = Oracle().eye_pass(make_scene("cornell", textures=None), abi.default_params(), reference_camera(24, 18), 0)
    order = np.lexsort((got["pix"][:, 1], got["pix"][:, 0]))          # canonical order: by pixel (stable within a pixel)
    worder = np.lexsort((whole["pix"][:, 1], whole["pix"][:, 0]))
    assert got["pos"].shape == whole["pos"].shape
    for k in ("pos", "n", "wgt", "r2", "pix"):
        np.testing.assert_array_equal(got[k][order], whole[k][worder])
