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
