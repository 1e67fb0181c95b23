#!/usr/bin/env python
"""bench.py — photons traced+gathered per second (BASELINE.json metric M1) on N B200s.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl b200|reference]
                  [--scaling weak|strong] [--update reference|progressive]

A step = one round of RayTracer::PhotonMap (Raytracer.h:226-246): the ranks trace the round's photons
(Light::emit + PhotonTrace), gather them into the replicated hit-point map, the per-round accumulators are
summed over ranks (NCCL, N > 1) and HitPoint::update runs.  Weak scaling (default workload): `onetime` photons
PER GPU per step; strong scaling (caustics-1080p, stress-4k: BASELINE configs 4 / 5): the step's `onetime`
photons are sharded over the GPUs.  The hit-point map (eye pass + hash grid) is built once before the timed
region, as in the reference where one eye pass serves cnt rounds; `e2e` times RayTracer.render() of the host
mirror (scene upload from pinned host memory, eye pass, grid build, cnt rounds, film resolve + read-back).

Besides `value` the line carries: `roofline` (the kernel class with the largest total time per step, from
serialised rounds timed with CUDA events between launches), `kernels` (every class the same way), `exact_mode`
(the same steps with the reference's arithmetic order), `multi_gpu_parity` (N > 1), `cpu_baseline`.

--impl reference times the CPU restatement of the reference (oracle/, OpenMP over all host cores, each thread
an independent photon tracer as in Raytracer.h:442-447) on a bounded sample of the same workload; the GPU is
not touched.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# name -> scene, film size, photons per light per round (BASELINE.json configs[1..4])
WORKLOADS = {
    "ref-main-320x240": dict(scene="reference", w=320, h=240, onetime=100_000),
    "cornell-1024x768": dict(scene="cornell", w=1024, h=768, onetime=1_000_000),
    "teapot-room-1024x768": dict(scene="reference", w=1024, h=768, onetime=2_000_000),
    "caustics-1080p": dict(scene="cornell", w=1920, h=1080, onetime=16_000_000),
    "stress-4k": dict(scene="stress", w=3840, h=2160, onetime=64_000_000),
}
DEFAULT_WORKLOAD = "teapot-room-1024x768"  # the scene the north-star target (>=1e9 photons/s on 8 GPUs) is quoted on


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "200"], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons, pw = [], [], set(), []
        for line in self.f.read().splitlines():
            c = [x.strip() for x in line.split(",")]
            if len(c) < 7:
                continue
            try:
                sm.append(float(c[0]))
                mx.append(float(c[1]))
                pw.append(float(c[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.f.name)
        if sm:
            out = {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons),
                   "power_w_max": max(pw), "samples": len(sm)}
        return out


def workload_config(name, n_hp, update):
    """`config` of the JSON line: what defines the workload, identical in both arms (b200 / reference).  What an
    arm does to run it (arithmetic modes, tracer, sample size, sharding) goes into `impl_config`."""
    wl = WORKLOADS[name]
    return {"workload": name, "scene": wl["scene"], "film": [wl["w"], wl["h"]], "hit_points": int(n_hp),
            "photons_per_round": wl["onetime"], "newton": "50 restarts x 10 iterations, AABB culling", "update_mode": update,
            "textures": "synthetic, reference sizes",
            "l2": "inputs larger than L2 (deposit list + hit-point map > 126 MB); new photons every step"}


def make_params(abi, fast=True):
    p = abi.default_params()
    p.newton_mode = abi.SPPM_NEWTON_FAST if fast else abi.SPPM_NEWTON_EXACT
    p.math_mode = abi.SPPM_MATH_FAST if fast else abi.SPPM_MATH_EXACT
    return p


def host_threads(orc) -> int:
    """All host cores this process may use.  torchrun exports OMP_NUM_THREADS=1, which would
    silently turn the CPU arm into a single-thread run: the oracle takes an explicit thread
    count (num_threads clause), so ask the OS instead of OpenMP."""
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    return max(n, orc.max_threads(), 1)


def reference_code_baseline(wl_name, scene, cam, threads, target_s):
    """The reference's OWN PhotonMap (its headers compiled against the OpenCV/FLANN shim,
    oracle/_ref/libref.so — built where /root/reference exists and shipped prebuilt) on `threads`
    independent replicas, as SPPMRayTracer::render runs it (Raytracer.h:429-447): onetime = 100
    photons per round as hard-coded (:219), the update sweep over every hit point after each
    round (:238-245).  Reported next to the oracle port, which is the faster of the two (it
    traces a whole sample per round) and therefore the baseline the ratio is taken against.
    None when the library is absent."""
    try:
        from oracle import ref as ref_mod
        if not ref_mod.available():
            return None
        r = ref_mod.Reference()
        if not r.can_time:
            return None
        h = r.scene(scene)
        o = r.time_photon_map(h, cam.pos, cam.look, cam.w, cam.h, 1_000_000, threads, target_s)
    except Exception as e:  # the checker must never take the benchmark down
        return {"unavailable": f"{type(e).__name__}: {e}"}
    return {"value": o["photons"] / max(o["map_s"], 1e-9), "unit": "photons/s", "cores": threads,
            "sample": f"{threads} replicas x {o['rounds']} rounds x 100 photons (onetime as shipped) of {wl_name}, "
                      f"{o['hit_points']} hit points per replica, update sweep per round, kd-tree restated in the FLANN shim; "
                      f"{o['map_s']:.1f} s timed, {o['setup_s']:.1f} s untimed eye pass + index build"}


def cpu_photon_baseline(wl, scene, prm, hp, target_s=12.0, cam=None):
    """The oracle's photon pass on all host cores (each thread an independent tracer over the
    shared hit-point map + exact kd-tree), sized to ~target_s seconds."""
    from oracle import Oracle
    orc = Oracle()
    threads = host_threads(orc)
    s, c = orc.photon_bench(scene, prm, hp, 1500, threads, pass_id=1000)
    rate = c["photons"] / max(s, 1e-6)
    per_thread = int(max(1500, min(4_000_000, rate * target_s / threads)))
    s, c = orc.photon_bench(scene, prm, hp, per_thread, threads, pass_id=1001)
    # the reference as shipped runs 4 OpenMP threads = 4 independent replicas (Raytracer.h:429-447)
    s4, c4 = orc.photon_bench(scene, prm, hp, max(1500, per_thread // 3), min(4, threads), pass_id=1002)
    ref_code = reference_code_baseline(wl, scene, cam, threads, 4.0) if cam is not None else None
    out = dict(value=c["photons"] / s, unit="photons/s", cores=threads, kind="port",
               sample=f"{threads} threads x {per_thread} photons of {wl} (hit-point map {hp['pos'].shape[0]} pts, exact kd-tree), "
                      f"{s:.1f} s wall; {c['solves']} Bezier solves",
               as_shipped_4_threads=c4["photons"] / s4, reference_code=ref_code)
    return _faster_of(out), s, c


def _faster_of(cpu):
    """The CPU figure the GPU is compared with is the FASTER of the two CPU implementations: the
    oracle port (one long round per sample) or the reference's own code (100-photon rounds, a
    sweep over all hit points after each).  Both stay in the record."""
    rc = cpu.get("reference_code")
    cpu["port_value"] = cpu["value"]
    if rc and rc.get("value", 0) > cpu["value"]:
        cpu.update(value=rc["value"], kind="reference", sample=rc["sample"])
    return cpu


def run_reference(args):
    """CPU arm: no GPU.  Eye pass + kd-tree by the oracle (untimed setup, as for the GPU arm),
    then K steps, each a bounded photon sample on all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import Oracle
    from raytrace3_b200 import abi
    from raytrace3_b200.scene import make_scene, reference_camera
    wl = WORKLOADS[args.workload]
    orc = Oracle()
    prm = make_params(abi, fast=False)
    scene = make_scene(wl["scene"])
    cam = reference_camera(wl["w"], wl["h"])
    t0 = time.time()
    hp = orc.eye_pass(scene, prm, cam, 0)
    t_eye = time.time() - t0
    threads = host_threads(orc)
    s, c = orc.photon_bench(scene, prm, hp, 1000, threads, pass_id=999)
    rate = c["photons"] / max(s, 1e-6)
    budget = 150.0 / max(1, args.steps + args.warmup)  # whole run within a few minutes
    step_s = float(os.environ.get("SPPM_BENCH_SAMPLE_S", "10.0"))  # seconds of CPU work per step (tests shrink it)
    per_thread = int(max(500, min(2_000_000, rate * min(budget, step_s) / threads)))
    times, photons, solves = [], 0, 0
    for i in range(args.warmup + args.steps):
        s, c = orc.photon_bench(scene, prm, hp, per_thread, threads, pass_id=i)
        if i >= args.warmup:
            times.append(s)
            photons += c["photons"]
            solves += c["solves"]
    total = sum(times)
    value = photons / total
    sample = (f"{threads} threads x {per_thread} photons per step of {args.workload} "
              f"(CPU eye pass {t_eye:.1f} s untimed, {hp['pos'].shape[0]} hit points, exact kd-tree)")
    ref_code = reference_code_baseline(args.workload, scene, cam, threads, min(step_s, 5.0))
    cpu = _faster_of({"value": value, "unit": "photons/s", "cores": threads, "kind": "port", "sample": sample, "reference_code": ref_code})
    if cpu["kind"] == "reference":  # the reference's own code beat the port on this workload: it is the line's value
        value, total = cpu["value"], (photons / cpu["value"])
    line = {
        "impl": "reference", "metric": "photons traced+gathered/sec", "value": value, "unit": "photons/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / max(1, args.steps),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args.workload, hp["pos"].shape[0], WORKLOAD_DEFAULTS.get(args.workload, {}).get("update", "reference")),
        "impl_config": {"photons_per_step": threads * per_thread, "arithmetic": "exact order (oracle = the reference operation by operation)",
                        "update": "none inside the sample (the reference's update is a no-op on fresh hit points, Raytracer.h:70)"},
        "cpu_baseline": cpu,
        "e2e": {"value": value, "unit": "photons/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "bezier_isects_per_s": solves / total,
    }
    _emit(line)


# BASELINE.json: config 4 shards 16 M photons/pass over the GPUs (strong), config 5 runs radius-shrinking passes
WORKLOAD_DEFAULTS = {
    "caustics-1080p": dict(scaling="strong", update="reference"),
    "stress-4k": dict(scaling="strong", update="progressive"),
}


def csrc_digest() -> str:
    """Digest of the kernel sources: ncu-derived numbers in profiles/ncu_traffic.json are only quoted when they
    were captured from this very source."""
    import hashlib
    h = hashlib.sha256()
    d = os.path.join(ROOT, "raytrace3_b200", "csrc")
    for f in sorted(os.listdir(d)):
        if f.endswith(".cuh"):  # the kernel bodies (the .cu files hold launchers and plumbing)
            h.update(open(os.path.join(d, f), "rb").read())
    return h.hexdigest()[:16]


def ncu_traffic(kernel: str, workload: str):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel` from the committed `ncu --set full`
    capture (profiles/ncu_traffic.json, written by tools/ncu_traffic.py from a capture of THIS kernel source), or
    None: a number captured from another version of the kernels is not quoted."""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            t = json.load(f)
        if t.get("csrc_digest") != csrc_digest():
            return None, None
        e = t.get("kernels", {}).get(f"{kernel}@{workload}")
        return (float(e["dram_bytes"]), e.get("source")) if e else (None, None)
    except Exception:
        return None, None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--newton", default="fast", choices=["fast", "exact"])
    ap.add_argument("--trace", default="wavefront", choices=["wavefront", "persistent"])
    ap.add_argument("--gather", default="filtered", choices=["filtered", "exact"],
                    help="gather predicate: packed-FP32 interval filter + FP64 fallback (same newN), or FP64 on every pair")
    ap.add_argument("--update", default=None, choices=["reference", "progressive"],
                    help="HitPoint::update as shipped (a no-op on fresh hit points, Q1) or the radius-shrinking rule "
                         "(default: progressive for stress-4k, as BASELINE config 5 states; reference otherwise)")
    ap.add_argument("--scaling", default=None, choices=["weak", "strong"],
                    help="weak: `onetime` photons per GPU and step; strong: the step's `onetime` photons are sharded over the "
                         "GPUs (default for caustics-1080p and stress-4k, as BASELINE configs 4 / 5 state)")
    ap.add_argument("--count-pairs", action="store_true", help="keep the diagnostic cand_pairs counter on in the timed loop")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-exact-mode", action="store_true", help="skip the EXACT-arithmetic leg (exact_mode in the line)")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--e2e-rounds", type=int, default=10, help="PhotonMap rounds per RayTracer::render() in the e2e leg "
                    "(the reference runs 10000 rounds per eye pass, Raytracer.h:384)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    assert world == args.gpus or world == 1 and args.gpus == 1, f"launch with torchrun --nproc-per-node {args.gpus}"

    import torch
    from raytrace3_b200 import Context, RayTracer, abi, photon_range
    from raytrace3_b200.scene import make_scene, reference_camera

    dist = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    wl = WORKLOADS[args.workload]
    wd = WORKLOAD_DEFAULTS.get(args.workload, {})
    scaling = args.scaling or wd.get("scaling", "weak")
    update = args.update or wd.get("update", "reference")

    def params_for(fast):
        p = make_params(abi, fast=fast)
        p.trace_mode = abi.SPPM_TRACE_PERSISTENT if args.trace == "persistent" else abi.SPPM_TRACE_WAVEFRONT
        p.gather_mode = abi.SPPM_GATHER_EXACT if args.gather == "exact" else abi.SPPM_GATHER_FILTERED
        p.update_mode = abi.SPPM_UPDATE_PROGRESSIVE if update == "progressive" else abi.SPPM_UPDATE_REFERENCE
        p.count_pairs = 1 if args.count_pairs else 0   # diagnostic pair counter: off in the timed loop, see below
        return p

    prm = params_for(args.newton == "fast")
    ctx = Context(local_rank, prm)
    scene = make_scene(wl["scene"])
    cam = reference_camera(wl["w"], wl["h"])
    onetime = wl["onetime"]
    n_light = len(scene.plvec)
    # photons this rank traces per step, and the whole job's
    if scaling == "weak":
        my_begin, my_count, step_photons = rank * onetime, onetime, onetime * n_light * world
    else:
        my_begin, my_count = photon_range(onetime, rank, world)
        step_photons = onetime * n_light

    def barrier():
        ctx.synchronize()
        if dist is not None:
            dist.barrier()

    # ---- setup (untimed): scene, eye pass, hash grid, communicator
    ctx.upload_scene(scene)
    n_hp = ctx.eye_pass(cam, 0)
    ctx.build_index()
    setup = ctx.counters()
    if world > 1:
        uid = [ctx.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        ctx.comm_init_rank(rank, world, uid[0])

    def step(i):
        # round i: rank r traces its photon range of the round (ids are global: any sharding traces the same set)
        ctx.photon_pass(my_begin, my_count, 1 + i)
        if world > 1:
            ctx.allreduce()
        ctx.update()

    def timed(first, k):
        barrier()
        ctx.timer_record(0)
        for i in range(k):
            step(first + i)
        ctx.timer_record(1)
        t = ctx.timer_elapsed_ms()
        barrier()
        if dist is not None:
            tt = torch.tensor([t], dtype=torch.float64, device=f"cuda:{local_rank}")
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            t = float(tt.item())
        return t

    # set-up, untimed: one round per pipeline lane so that every lane has allocated its buffers (a lane's first
    # batch pays cudaMalloc); then the W warm-up steps the caller asked for
    next_round = 0
    n_prime = 8 if my_count * n_light <= (1 << 19) else 4   # the library rotates small batches over more lanes (ctx.h)
    for i in range(n_prime):
        step(next_round + i)
    next_round += n_prime
    for i in range(args.warmup):
        step(next_round + i)
    next_round += args.warmup
    barrier()
    ctx.reset_counters()
    launches0 = ctx.counters()["launches"]
    sampler = ClockSampler(local_rank) if rank == 0 else None
    ms = timed(next_round, args.steps)
    next_round += args.steps
    clocks = sampler.stop() if sampler else None
    cnt = ctx.counters()
    steps = args.steps
    value = step_photons * steps / (ms * 1e-3)

    # ---- the same steps in EXACT arithmetic (Newton as written in Bezier.h:112-159 without FMA, sampling through
    # acos / sin / cos as Vec3.h:90-115): what bit-exactness with the reference costs
    exact_mode = None
    if args.newton == "fast" and not args.no_exact_mode:
        ctx.set_params(params_for(False))
        k = max(2, min(steps, 10))
        for i in range(3):
            step(next_round + i)
        ms_x = timed(next_round + 3, k)
        next_round += 3 + k
        exact_mode = {"value": step_photons * k / (ms_x * 1e-3), "unit": "photons/s", "ms_per_step": ms_x / k, "steps": k,
                      "newton": "50x10 exact (Bezier.h:112-159 operation order, no FMA; bit-identical to the reference's own code)",
                      "math": "exact (Vec3.h:90-115 through acos/sin/cos)"}
        ctx.set_params(prm)

    # ---- work statistics (cand_pairs feeds the SURVEY 8d flop count): two untimed rounds with the diagnostic
    # counter on -- the timed loop above runs with it off unless --count-pairs
    prm.count_pairs = 1
    ctx.set_params(prm)
    ctx.reset_counters()
    for i in range(2):
        step(next_round + i)
    next_round += 2
    barrier()
    stat = ctx.counters()
    cnt["cand_pairs"] = stat["cand_pairs"] * steps // 2
    prm.count_pairs = 1 if args.count_pairs else 0
    ctx.set_params(prm)

    # ---- serialised rounds: every batch alone on the GPU with a CUDA event between its launches
    # (sppm_set_profiling), so each kernel class gets the time an ncu launch list gives it
    n_serial = 4 if onetime <= 4_000_000 else 2
    ctx.reset_counters()
    ctx.set_profiling(True)
    for i in range(n_serial):
        step(next_round + i)
        barrier()
    next_round += n_serial
    ser = ctx.counters()
    ctx.set_profiling(False)
    ser_ms = {k: ser[f"ms_{k}"] / n_serial for k in ("step", "newton", "spill", "bin_scan", "bin_scatter", "gather_prep",
                                                     "gather_kernel", "update", "allreduce")}
    ser_round = sum(ser_ms.values())
    ser_work = {k: ser[k] / n_serial for k in ("isect_queries", "patch_solves", "patch_culled", "deposits_kept", "acc_pairs",
                                                "spilled_photons", "photons")}

    # ---- e2e: RayTracer::render() of the host mirror (the call a user makes) with pinned host buffers
    e2e_ms, e2e_phases = [], None
    pins = []
    for tex in {id(o.texture): o.texture for o in scene.objvec if o.texture is not None}.values():
        buf = torch.empty(tex.map.shape, dtype=torch.uint8, pin_memory=True)
        arr = buf.numpy()
        arr[...] = tex.map
        tex.map = arr
        pins.append(buf)
    film_pin = torch.zeros((wl["h"], wl["w"], 3), dtype=torch.float64, pin_memory=True)
    cam.pic = film_pin.numpy()
    h2d = scene.texture_bytes() + sum(o.P.nbytes for o in scene.objvec if o.kind == 2) + 328 * len(scene.objvec) + 256
    d2h = cam.pic.nbytes + 8 * 3
    e2e_onetime = onetime * world if scaling == "weak" else onetime     # photons per round over all ranks
    rt = RayTracer(ctx, rank=rank, world=world)
    rt.camera, rt.scene, rt.cnt, rt.onetime = cam, scene, args.e2e_rounds, e2e_onetime
    # two untimed renders first: a render brings fewer batches than the context has lanes, and every lane sizes its
    # buffers the first time it sees a batch of this size
    E2E_WARM = 2
    for i in range(args.e2e_steps + E2E_WARM if args.e2e_steps > 0 else 0):
        barrier()
        t0 = time.perf_counter()
        rt._scene_uploaded = None      # H2D every step: objects, patches, textures from pinned host memory
        rt.pass_base = 100 + i
        rt.render()                    # eye pass, grid build, cnt rounds, resolve, D2H of the film into cam.pic
        barrier()
        t1 = time.perf_counter()
        if i >= E2E_WARM:
            e2e_ms.append((t1 - t0) * 1e3)
    e2e = statistics.median(e2e_ms) if e2e_ms else float('nan')
    if dist is not None:
        t = torch.tensor([e2e], dtype=torch.float64, device=f"cuda:{local_rank}")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e = float(t.item())
    e2e_value = e2e_onetime * n_light * args.e2e_rounds / (e2e * 1e-3)
    e2e_counters = ctx.counters()
    # the same call in the reference's OWN loop shape (Raytracer.h:218, 384: cnt = 10000 rounds of onetime = 100
    # photons per eye pass, i.e. an eye pass and a grid build per 1 M photons); one GPU only, reported beside e2e
    ref_shape = None
    if world == 1 and args.e2e_steps > 0 and update == "reference":
        try:
            rs_ms = []
            rt.cnt, rt.onetime = 10000, 100
            for i in range(E2E_WARM + 5):
                ctx.synchronize()
                t0 = time.perf_counter()
                rt._scene_uploaded = None
                rt.pass_base = 200 + i
                rt.render()
                t1 = time.perf_counter()
                if i >= E2E_WARM:
                    rs_ms.append((t1 - t0) * 1e3)
            m = statistics.median(rs_ms)
            ref_shape = {"value": 10000 * 100 * n_light / (m * 1e-3), "unit": "photons/s", "ms_per_render": m,
                         "what": "RayTracer.render() with the reference's literals cnt=10000, onetime=100: scene upload, eye pass, "
                                 "grid build, 1 M photons (the rounds fused into one pass), resolve, film read-back"}
        except Exception as e:  # a side figure: never costs the line
            ref_shape = {"error": str(e)[:200]}

    # ---- multi-GPU parity (world > 1): one small round sharded over the ranks + allreduce must leave every rank
    # with the same accumulators, equal to the same round traced by rank 0 alone
    mgp = None
    if world > 1:
        mgp = multi_gpu_parity(ctx, Context, abi, torch, dist, scene, cam, prm, rank, world, local_rank, photon_range)

    if rank != 0:
        if dist is not None:
            dist.barrier()
            dist.destroy_process_group()
        return

    # ---- rooflines.  Algorithmic work per launch: bytes and flops per unit x units per step,
    # divided by the launches per step; times from the serialised rounds above (CUDA events on the launching
    # stream, measured live in this run).
    peaks, peak_kind = measured_peaks()
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    fp64_peak = ctx.fp64_peak()
    per = lambda k: cnt[k] / steps
    dep_kept, cand, acc = per("deposits_kept"), per("cand_pairs"), per("acc_pairs")
    solves, culled, queries = per("patch_solves"), per("patch_culled"), per("isect_queries")
    n_batches = max(1, -(-my_count * n_light // int(os.environ.get("SPPM_PHOTON_BATCH", 1 << 23))))  # abi.cu PHOTON_BATCH
    n_bounce = prm.max_dep + 2
    has_bezier = any(o.kind == 2 for o in scene.objvec)
    wave = prm.trace_mode != abi.SPPM_TRACE_PERSISTENT
    gather_name = "gather_kernel" if args.gather == "exact" else "gather_filtered_kernel"
    K = {}

    def add(name, ms_k, launches, bytes_, flops, bound):
        if ms_k <= 0:
            return
        gbs = bytes_ / (ms_k * 1e-3) / 1e9
        tf = flops / (ms_k * 1e-3) / 1e12
        K[name] = {"ms_per_step": ms_k, "launches_per_step": launches, "ms_per_launch": ms_k / launches,
                   "share_of_round": ms_k / ser_round if ser_round > 0 else None, "bound": bound,
                   "algorithmic_bytes_per_step": bytes_, "algorithmic_flops_per_step": flops,
                   "GBps": gbs, "hbm_frac": gbs / hbm_peak, "fp64_tflops": tf,
                   "fp64_frac": tf / fp64_peak if fp64_peak > 0 else None}

    executed = max(solves - culled, 0.0)
    if wave:
        # step kernel: 100 B of queue state read + 100 B written per live photon and bounce, 88 B per kept deposit
        # (80 B record + key + rank), 40 B per (ray, patch) item (slot, patch written; t, u, v read back)
        add("photon_step_kernel", ser_ms["step"], n_bounce * n_batches, 200.0 * queries + 88.0 * dep_kept + 40.0 * solves,
            140.0 * queries, "latency (issue slots / long scoreboard), HBM second")
        if has_bezier:
            add("newton_items_kernel", ser_ms["newton"], (n_bounce - 1) * n_batches, 40.0 * solves, 175000.0 * executed, "fp64")
    else:
        add("photon_trace_kernel", ser_ms["spill"] + ser_ms["step"], n_batches, 88.0 * dep_kept, 175000.0 * executed + 140.0 * queries, "fp64 / latency")
    add("deposit_binning (cub scan + dep_scatter_kernel)", ser_ms["bin_scan"] + ser_ms["bin_scatter"], 2 * n_batches,
        168.0 * dep_kept + 8.0 * n_hp / 4, 0.0, "hbm")
    add(gather_name, ser_ms["gather_kernel"], n_batches, 136.0 * n_hp + 72.0 * dep_kept, 13.0 * cand + 9.0 * acc,
        "instruction issue (packed FP32 classification), HBM is the stated roofline")
    add("update_kernel", ser_ms["update"], 1, 136.0 * n_hp, 6.0 * n_hp, "hbm")
    dom = max(K, key=lambda k: K[k]["ms_per_step"])
    d = K[dom]
    traffic, traffic_src = ncu_traffic(dom, args.workload) if world == 1 else (None, None)
    in_loop = {"trace+binning": cnt["ms_trace"] / steps, "gather(prep+kernel)": cnt["ms_gather"] / steps}
    roofline = {"kernel": dom, "bound": "hbm", "achieved": d["GBps"],
                "peak": hbm_peak, "unit": "GB/s", "frac": d["hbm_frac"], "traffic": traffic, "traffic_source": traffic_src,
                "algorithmic_bytes": d["algorithmic_bytes_per_step"] / d["launches_per_step"],
                "ms_per_launch": d["ms_per_launch"], "launches_per_step": d["launches_per_step"],
                "share_of_round": d["share_of_round"],
                "peak_source": f"MEASURED_PEAKS.json hbm_gbs ({peak_kind}, sustained copy)",
                "timing": "CUDA events on the launching stream around every launch of the class, in serialised rounds of this run "
                          "(sppm_set_profiling: a batch alone on the GPU, as ncu's launch list sees it); "
                          "in the pipelined timed loop the lanes' spans overlap (in_pipeline_span_ms) and do not add up to ms_per_step",
                "in_pipeline_span_ms": in_loop, "serialized_round_ms": ser_round,
                "fp64_tflops": d["fp64_tflops"], "fp64_peak_tflops": fp64_peak, "fp64_frac": d["fp64_frac"],
                "binding_resource": d["bound"],
                "dominant_by": "largest total time per step among the kernel classes (serialised)"}

    # ---- Bezier ray-patch intersections per second (metric M2), device-resident microbench
    m2 = None
    if has_bezier:
        npairs = 1 << 18
        ms_rep = ctx.bezier_patch_bench(npairs, 3)
        nt = K.get("newton_items_kernel")
        m2 = {"value": npairs / (ms_rep * 1e-3), "unit": "ray-patch isects/s", "pairs": npairs, "ms": ms_rep,
              "fp64_tflops": 175000.0 * npairs / (ms_rep * 1e-3) / 1e12, "fp64_peak_tflops": fp64_peak,
              "in_render": {"executed_solves_per_step": executed, "culled_per_step": culled,
                            "isects_per_s": executed / (nt["ms_per_step"] * 1e-3) if nt else None,
                            "fp64_frac": nt["fp64_frac"] if nt else None,
                            "what": "solves that ran (patch_solves - patch_culled) / serialised newton_items_kernel time"}}

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        hp = ctx.export_hitpoints()
        cpu, _, _ = cpu_photon_baseline(args.workload, scene, make_params(abi, fast=False), hp, cam=cam)

    ph = onetime * n_light if scaling == "strong" else my_count * n_light
    line = {
        "metric": "photons traced+gathered/sec", "value": value, "unit": "photons/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / steps, "higher_is_better": True, "scaling": scaling, "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": workload_config(args.workload, n_hp, update),
        "impl_config": {"photons_per_step": step_photons, "photons_per_gpu_per_step": my_count * n_light,
                        "arithmetic": f"newton {args.newton}, math {'fast' if args.newton == 'fast' else 'exact'} (exact_mode: the same steps in the reference's order)",
                        "tracer": args.trace, "gather": args.gather,
                        "parallelism": f"photons x{world} ({scaling}), hit points replicated, allreduce of dN/dtao per step"},
        "e2e": {"value": e2e_value, "unit": "photons/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                "ms_per_step": e2e, "rounds_per_step": args.e2e_rounds,
                "phases_ms": {"eye": e2e_counters["ms_eye"], "build": e2e_counters["ms_build"], "resolve": e2e_counters["ms_resolve"]},
                "what": f"RayTracer.render() of the host mirror (raytrace3_b200.api, twin of include/raytracer/Raytracer.h) with "
                        f"cnt={args.e2e_rounds}, onetime={e2e_onetime}: scene + textures from pinned host memory, eye pass, grid build, "
                        f"PhotonMap (rounds fused into one photon pass while the update is the reference's), resolve, film "
                        f"read-back into pinned camera.pic (the reference runs cnt=10000 rounds of 100 photons per eye pass: "
                        f"reference_loop_shape)",
                "reference_loop_shape": ref_shape},
        "gpu_launches": int(cnt["launches"] - launches0),
        "clocks": clocks,
        "roofline": roofline,
        "kernels": K,
        "exact_mode": exact_mode,
        "multi_gpu_parity": mgp,
        "allreduce_ms_per_step": cnt["ms_allreduce"] / steps if world > 1 else None,
        "cpu_baseline": cpu,
        "bezier_isects": m2,
        "per_photon": {"isect_queries": queries / ph, "deposits": per("deposits") / ph,
                       "deposits_kept": dep_kept / ph, "cand_pairs": cand / ph,
                       "patch_solves": solves / ph, "patch_solves_executed": executed / ph,
                       "spilled": per("spilled_photons") / ph},
        "setup_ms": {"eye": setup["ms_eye"], "build": setup["ms_build"], "eye_e2e": e2e_counters["ms_eye"], "build_e2e": e2e_counters["ms_build"]},
    }
    _emit(line)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def multi_gpu_parity(ctx, Context, abi, torch, dist, scene, cam, prm, rank, world, local_rank, photon_range):
    """One small round, sharded over the ranks (photon_range) + NCCL allreduce + update, against the same round
    traced by rank 0 alone in a second context: newN identical on every rank and equal to the one-GPU run's,
    tao within 1e-11 (summation order).  Returns "ok" or a description of the first difference."""
    import hashlib
    small = 200_000
    p = abi.SppmParams.from_buffer_copy(prm)
    p.count_pairs = 0
    ctx.set_params(p)
    ctx.eye_pass(cam, 0)                   # the same eye pass as the one-GPU run below (the e2e leg left pass 100 + i's map)
    ctx.build_index()                      # fresh accumulators
    b, c = photon_range(small, rank, world)
    ctx.photon_pass(b, c, 7777)
    ctx.allreduce()
    ctx.update()
    n = ctx.counters()["hit_points"]
    st = ctx.export_stats(n)
    count = st["newN"] + st["N"]      # PROGRESSIVE update folds newN into N (alpha = 0.7, truncated) and clears it
    digest = hashlib.sha256(count.tobytes()).digest()[:8]
    mine = torch.tensor(list(digest), dtype=torch.uint8, device=f"cuda:{local_rank}")
    allv = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(allv, mine)
    same_ranks = all(bool((v == allv[0]).all().item()) for v in allv)
    tsum = torch.tensor(st["tao"].sum(0), dtype=torch.float64, device=f"cuda:{local_rank}")
    tall = [torch.empty_like(tsum) for _ in range(world)]
    dist.all_gather(tall, tsum)
    out = "ok"
    if not same_ranks:
        out = "newN differs between ranks"
    if rank == 0:
        with Context(local_rank, p) as solo:
            solo.upload_scene(scene)
            m = solo.eye_pass(cam, 0)
            solo.build_index()
            solo.photon_pass(0, small, 7777)
            solo.update()
            ref = solo.export_stats(m)
        if m != n or not (np.array_equal(ref["newN"], st["newN"]) and np.array_equal(ref["N"], st["N"])):
            out = f"newN of the {world}-GPU round differs from the 1-GPU round ({int((ref['newN'] + ref['N'] != count).sum()) if m == n else 'hit-point counts differ'})"
        elif not np.allclose(ref["tao"], st["tao"], rtol=1e-11, atol=1e-300):
            out = "tao differs from the 1-GPU round beyond 1e-11"
        elif int(count.sum()) == 0:
            out = "empty round"
        else:
            rel = max(abs(float(t[k]) - float(tall[0][k])) / max(1e-300, abs(float(tall[0][k]))) for t in tall for k in range(3))
            if rel > 1e-12:
                out = f"tao sums differ between ranks by {rel:.2e}"
            else:
                out = {"status": "ok", "photons": small, "ranks": world, "newN_sum": int(count.sum()),
                       "what": "newN bit-identical on all ranks and equal to a 1-GPU run of the same round; tao <= 1e-11"}
    return out


_REAL_STDOUT = None


def _emit(line: dict):
    """The ONE JSON line of the contract goes to the real stdout; everything else libraries print
    while the benchmark runs (NCCL's version banner, build logs) was redirected to stderr."""
    os.write(_REAL_STDOUT if _REAL_STDOUT is not None else 1, (json.dumps(line) + "\n").encode())


def _quiet_stdout():
    """Keep stdout clean for the driver: fd 1 -> stderr until the result line is written."""
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)


if __name__ == "__main__":
    _quiet_stdout()
    main()
