"""Multi-GPU correctness check (run with torchrun, one rank per GPU; also checks the sharded eye pass):
  python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/check_multigpu.py
Every rank traces its photon_range() slice of two rounds, the per-round accumulators are summed
with sppm_allreduce (NCCL), and rank 0 compares the result with a single-GPU run of the whole
range: newN must be bit-exact, tao equal to rounding; all ranks must end bit-identical."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from raytrace3_b200 import Context, abi, photon_range, row_range  # noqa: E402
from raytrace3_b200.scene import make_scene, reference_camera  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    prm = abi.default_params()
    prm.newton_mode = abi.SPPM_NEWTON_FAST
    ctx = Context(local, prm)
    sc = make_scene("reference", textures=None)
    cam = reference_camera(256, 192)
    ctx.upload_scene(sc)
    uid = [ctx.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    ctx.comm_init_rank(rank, world, uid[0])
    # eye-pass sharding: every rank traces its rows, the shares are exchanged over
    # NCCL, and the map must equal the one-GPU eye pass hit point by hit point
    r0, r1 = row_range(cam.h, rank, world)
    mine = ctx.eye_pass_rows(cam, 0, r0, r1)
    n = ctx.allgather_hitpoints()
    assert n >= mine > 0
    ctx.build_index()
    sharded = ctx.export_hitpoints()
    if rank == 0:
        one = Context(local, prm)
        one.upload_scene(sc)
        assert one.eye_pass(cam, 0) == n, "sharded eye pass lost or duplicated hit points"
        whole = one.export_hitpoints()
        for k in ("pos", "n", "wgt", "r2", "pix", "node"):
            np.testing.assert_array_equal(sharded[k], whole[k])
        one.close()
    onetime = 300_001
    for rnd in range(2):
        b, c = photon_range(onetime, rank, world)
        ctx.photon_pass(b, c, 10 + rnd)
        ctx.allreduce()
        ctx.update()
    got = ctx.export_stats(n)
    # all ranks identical
    t = torch.from_numpy(got["newN"].astype(np.int64)).cuda()
    lo, hi = t.clone(), t.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    assert torch.equal(lo, hi), "ranks disagree on newN"
    tt = torch.from_numpy(got["tao"]).cuda()
    lo, hi = tt.clone(), tt.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    assert torch.equal(lo, hi), "ranks disagree on tao (replicas must stay bit-identical)"
    if rank == 0:
        ref = Context(local, prm)
        ref.upload_scene(sc)
        assert ref.eye_pass(cam, 0) == n
        ref.build_index()
        for rnd in range(2):
            ref.photon_pass(0, onetime, 10 + rnd)
            ref.update()
        want = ref.export_stats(n)
        assert want["newN"].sum() > 100000
        np.testing.assert_array_equal(got["newN"], want["newN"])
        np.testing.assert_allclose(got["tao"], want["tao"], rtol=1e-12, atol=0)
        print(f"multi-GPU check ok: world={world}, hit points={n}, newN sum={int(want['newN'].sum())}")
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
