"""tests/device_host — TEST INFRASTRUCTURE: the product's device headers compiled for the CPU.

`DeviceHost()` builds (g++, a few seconds) and loads tests/device_host/_build/libdevice_host.so:
raytrace3_b200/csrc/{device_math.cuh, scene.cuh, scene_prep.h, trace_kernels.cuh} through
cuda_host_shim.h, with a warp emulated by 32 fibers and kernels launched with the product's block sizes (or one-warp blocks).  It lets the `-m "not gpu"` suite hold the very source lines the
sm_100a kernels are compiled from to the oracle.  Nothing in raytrace3_b200/ imports this."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from raytrace3_b200 import abi
from raytrace3_b200.api import hit_dtype

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(os.path.dirname(_HERE))
_CSRC = os.path.join(_ROOT, "raytrace3_b200", "csrc")
LIB_PATH = os.path.join(_HERE, "_build", "libdevice_host.so")
SOURCES = [os.path.join(_HERE, f) for f in ("device_host.cpp", "cuda_host_shim.h")] + [
    os.path.join(_CSRC, f) for f in ("scene.cuh", "device_math.cuh", "scene_prep.h", "trace_kernels.cuh", "wavefront_kernels.cuh", "gather_kernels.cuh", "image_kernels.cuh", "ctx.h")] + [
    os.path.join(_ROOT, "include", "sppm_b200.h")]
# -ffp-contract=off: no FMA contraction, the library's -fmad=false (explicit fma() calls stay fused)
# -fno-extern-tls-init: the kernels' block-scope `extern __shared__` (= thread_local here) array has no dynamic initialiser
CXXFLAGS = ["-O2", "-std=c++17", "-ffp-contract=off", "-fno-extern-tls-init", "-fPIC", "-shared", "-pthread"]

VP, I32, I64, U64, U32, DBL = C.c_void_p, C.c_int32, C.c_int64, C.c_uint64, C.c_uint32, C.c_double

# ctx.h HpRaw / Deposit
hpraw_dtype = np.dtype([("pos", "f8", 3), ("n", "f8", 3), ("wgt", "f8", 3), ("r2", "f8"), ("row", "i4"), ("col", "i4"),
                        ("okey", "u8")])
deposit_dtype = np.dtype([("x", "f8", 3), ("n", "f8", 3), ("flux", "f8", 3), ("meta", "u8")])
COUNTERS = ("photons", "isect_queries", "patch_solves", "deposits", "deposits_kept", "cand_pairs", "acc_pairs", "overflow",
            "patch_culled")


class Grid(C.Structure):  # ctx.h Grid
    _fields_ = [("org", DBL * 3), ("inv_cell", DBL), ("dim", C.c_int * 3), ("n_cells", I64), ("keep_lo", DBL * 3),
                ("keep_hi", DBL * 3)]


def make_grid(pos, search_r2) -> Grid:
    """The hash grid build_index (gather.cu) lays over the hit points: cell edge sqrt(searchR2) (1 + 1e-9),
    origin at their minimum, deposits kept within one cell edge of their bounding box (no 2^26-cell cap here)."""
    cell = float(np.sqrt(search_r2)) * (1.0 + 1e-9)
    lo, hi = pos.min(0), pos.max(0)
    dim = [int(np.floor((hi[k] - lo[k]) / cell) + 1) for k in range(3)]
    g = Grid()
    for k in range(3):
        g.org[k], g.dim[k], g.keep_lo[k], g.keep_hi[k] = lo[k], dim[k], lo[k] - cell, hi[k] + cell
    g.inv_cell = 1.0 / cell
    g.n_cells = dim[0] * dim[1] * dim[2]
    return g


def cuda_include_dir() -> str:
    for d in (os.environ.get("CUDA_HOME"), "/usr/local/cuda"):
        if d and os.path.exists(os.path.join(d, "include", "cuda_runtime.h")):
            return os.path.join(d, "include")
    raise RuntimeError("cuda_runtime.h not found (CUDA_HOME)")


# SPPM_DEVICE_HOST_SANITIZE=1: the same sources with -fsanitize=address,undefined (a memcheck of the kernel
# bodies on the CPU; run python with LD_PRELOAD=libasan.so, see test_device_kernels_are_clean_under_asan_ubsan)
SANITIZE = bool(os.environ.get("SPPM_DEVICE_HOST_SANITIZE"))
if SANITIZE:
    LIB_PATH = os.path.join(_HERE, "_build", "libdevice_host_asan.so")
    CXXFLAGS = ["-O1", "-g", "-fsanitize=address,undefined", "-fno-omit-frame-pointer"] + CXXFLAGS[1:]


# SPPM_DEVICE_HOST_EXTRA="-DSPPM_UNIT_HP=32 ...": the library's compile-time knobs (SPPM_NVCC_EXTRA of raytrace3_b200/build.py)
# for the emulation, in a library of its own (tools/gather_work_model.py compares them)
EXTRA = os.environ.get("SPPM_DEVICE_HOST_EXTRA", "").split()
if EXTRA:
    import hashlib
    LIB_PATH = LIB_PATH.replace(".so", "_" + hashlib.md5(" ".join(EXTRA).encode()).hexdigest()[:8] + ".so")
    CXXFLAGS = CXXFLAGS + EXTRA


def build(force: bool = False) -> str:
    stale = force or not os.path.exists(LIB_PATH) or any(
        os.path.getmtime(s) > os.path.getmtime(LIB_PATH) for s in SOURCES)
    if stale:
        os.makedirs(os.path.dirname(LIB_PATH), exist_ok=True)
        tmp = LIB_PATH + ".%d.tmp" % os.getpid()
        cmd = ["g++"] + CXXFLAGS + ["-I", cuda_include_dir(), "-I", os.path.join(_ROOT, "include"), "-o", tmp, SOURCES[0]]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("device_host build failed:\n" + r.stdout + r.stderr)
        os.replace(tmp, LIB_PATH)
    return LIB_PATH


def _p(a):
    return None if a is None else a.ctypes.data


def _f64(a, shape=None):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a.reshape(shape) if shape is not None else a


class HostScene:
    """A scene flattened by the product's prep_scene (scene_prep.h), viewed as the kernels view it."""

    def __init__(self, dh: "DeviceHost", scene):
        self.dh = dh
        self.desc = scene.flatten() if hasattr(scene, "flatten") else scene
        err = C.create_string_buffer(256)
        self.h = dh.lib.dh_scene_create(C.byref(self.desc), err, 256)
        if not self.h:
            raise ValueError(err.value.decode())
        self.n_patches = dh.lib.dh_scene_n_patches(self.h)
        self.n_objects = self.desc.n_objects

    def close(self):
        if self.h:
            self.dh.lib.dh_scene_free(self.h)
            self.h = None

    def __del__(self):
        self.close()

    def tables(self):
        n, m = self.n_patches, self.n_objects
        box, dop, sub, A, obox = np.empty((n, 6)), np.empty((n, 26)), np.empty((n, 16, 6)), np.empty((n, 48)), np.empty((m, 6))
        self.dh.lib.dh_scene_tables(self.h, _p(box), _p(dop), _p(sub), _p(A), _p(obox))
        return dict(box=box, dop=dop, sub=sub, A=A, objbox=obox)

    def material(self, i):
        fl, w = np.zeros(3, dtype=np.int32), np.zeros(4)
        self.dh.lib.dh_scene_material(self.h, i, _p(fl), _p(w))
        return fl.astype(bool), w

    @property
    def culled(self):
        return self.dh.lib.dh_scene_culled(self.h)

    def analytic_nearest(self, rays6):
        rays6 = _f64(rays6, (-1, 6))
        n = rays6.shape[0]
        obj, t, inside = np.empty(n, dtype=np.int32), np.empty(n), np.empty(n, dtype=np.int32)
        self.dh.lib.dh_analytic_nearest(self.h, n, _p(rays6), _p(obj), _p(t), _p(inside))
        return obj, t, inside

    def newton_cull(self, rays6, patch):
        rays6 = _f64(rays6, (-1, 6))
        patch = np.ascontiguousarray(patch, dtype=np.int32)
        n = rays6.shape[0]
        dop, sub = np.empty(n, dtype=np.int32), np.empty(n, dtype=np.int32)
        self.dh.lib.dh_newton_cull(self.h, n, _p(rays6), _p(patch), _p(dop), _p(sub))
        return dop.astype(bool), sub.astype(bool)

    def near_collide(self, prm, rays6, ids=None, pass_id=0, node=0, cull=1):
        """Scene::NearCollide through intersect_kernel -> SppmHit records."""
        return self.intersect_kernel(prm, rays6, ids, pass_id, node, cull)[0]

    # ---- the kernels of trace_kernels.cuh, launched as one-warp blocks
    def eye_pass_kernel(self, prm, cam, pass_id=0, row_begin=0, row_end=None, cap=None):
        """eye_pass_kernel -> hit points in the canonical (pixel, node) order + counters."""
        cd = cam.desc() if hasattr(cam, "desc") else cam
        row_end = cd.h if row_end is None else row_end
        if cap is None:
            cap = cd.w * cd.h * 4 + 1024
        out = np.zeros(cap, dtype=hpraw_dtype)
        cnt = np.zeros(len(COUNTERS), dtype=np.uint64)
        n = self.dh.lib.dh_eye_pass_kernel(self.h, C.byref(prm), C.byref(cd), pass_id, row_begin, row_end, cap, _p(out), _p(cnt))
        if n > cap:   # as sppm_eye_pass_rows does: the counter holds the true count, trace again with room for it
            return self.eye_pass_kernel(prm, cam, pass_id, row_begin, row_end, cap=n)
        out = out[:n]
        return out[np.argsort(out["okey"], kind="stable")], dict(zip(COUNTERS, cnt.tolist()))

    def photon_trace_kernel(self, prm, begin, count, pass_id, blocks=24):
        """photon_trace_kernel (persistent warps), every diffuse deposit kept -> deposits sorted by
        (photon id, bounce) + counters."""
        cap = int(count * max(1, self.desc.n_lights) * (prm.max_dep + 2))
        out = np.zeros(cap, dtype=deposit_dtype)
        cnt = np.zeros(len(COUNTERS), dtype=np.uint64)
        n = self.dh.lib.dh_photon_trace_kernel(self.h, C.byref(prm), begin, count, pass_id, blocks, cap, _p(out), _p(cnt))
        assert n <= cap
        out = out[:n]
        photon = out["meta"] & np.uint64((1 << 56) - 1)
        bounce = (out["meta"] >> np.uint64(56)).astype(np.int32)
        order = np.lexsort((bounce, photon))
        return dict(x=out["x"][order], n=out["n"][order], flux=out["flux"][order], photon=photon[order],
                    bounce=bounce[order]), dict(zip(COUNTERS, cnt.tolist()))

    def photon_trace_wavefront(self, prm, begin, count, pass_id, step_blocks=24, newton_blocks=24, item_cap=0, tail_bounce=-1):
        """The bounce-synchronous pass (photon_step_kernel + newton_items_kernel per bounce, the
        product's wf_layout / wf_step_args / wf_newton_args) -> deposits sorted by (photon id, bounce),
        counters, and the queue counts [bounce][region]."""
        cap = int(count * max(1, self.desc.n_lights) * (prm.max_dep + 2))
        out = np.zeros(cap, dtype=deposit_dtype)
        cnt = np.zeros(len(COUNTERS), dtype=np.uint64)
        qc = np.zeros((prm.max_dep + 3, 4), dtype=np.uint32)
        n = self.dh.lib.dh_photon_trace_wavefront(self.h, C.byref(prm), begin, count, pass_id, step_blocks, newton_blocks,
                                                  cap, _p(out), _p(cnt), _p(qc), item_cap, tail_bounce)
        assert 0 <= n <= cap
        out = out[:n]
        photon = out["meta"] & np.uint64((1 << 56) - 1)
        bounce = (out["meta"] >> np.uint64(56)).astype(np.int32)
        order = np.lexsort((bounce, photon))
        return dict(x=out["x"][order], n=out["n"][order], flux=out["flux"][order], photon=photon[order],
                    bounce=bounce[order]), dict(zip(COUNTERS, cnt.tolist())), qc

    def intersect_kernel(self, prm, rays6, ids=None, pass_id=0, node=0, cull=1):
        rays6 = _f64(rays6, (-1, 6))
        n = rays6.shape[0]
        out = np.zeros(n, dtype=hit_dtype)
        if ids is not None:
            ids = np.ascontiguousarray(ids, dtype=np.uint64)
        cnt = np.zeros(len(COUNTERS), dtype=np.uint64)
        self.dh.lib.dh_intersect_kernel(self.h, C.byref(prm), cull, n, _p(rays6), _p(ids), pass_id, node, _p(out), _p(cnt))
        return out, dict(zip(COUNTERS, cnt.tolist()))

    def patch_batch_kernel(self, prm, rays6, patch, seeds=None, cull=0, blocks=16):
        rays6 = _f64(rays6, (-1, 6))
        n = rays6.shape[0]
        patch = np.ascontiguousarray(patch, dtype=np.int32)
        if seeds is not None:
            seeds = _f64(seeds, (n, prm.newton_restarts, 2))
        out = np.empty((n, 3))
        self.dh.lib.dh_patch_batch_kernel(self.h, C.byref(prm), cull, n, _p(rays6), _p(patch), _p(seeds), blocks, _p(out))
        return out


class DeviceHost:
    def __init__(self):
        self.lib = L = C.CDLL(build())
        L.dh_draw2.argtypes = [U64, U64, U64, U32, U32, U32, U32, VP]
        L.dh_unit_rand_vec.argtypes = [U64, U64, I64, VP, U32, VP]
        L.dh_aabb_intersect.argtypes = [VP, I64, VP, VP]
        L.dh_bezier_eval.argtypes = [C.c_int, VP, I64, VP, VP]
        L.dh_newton.argtypes = [C.c_int, VP, I64, VP, VP, C.c_int, C.c_int, VP]
        L.dh_color_uv.argtypes = [C.c_int, C.c_int, VP, I64, VP, VP, C.c_int]
        L.dh_unit_rand_refl.argtypes = [I64, VP, VP, C.c_int, VP]
        L.dh_refracted.argtypes = [I64, VP, VP, VP, VP]
        L.dh_scene_create.argtypes = [C.POINTER(abi.SppmSceneDesc), C.c_char_p, C.c_int]
        L.dh_scene_create.restype = VP
        L.dh_scene_free.argtypes = [VP]
        L.dh_scene_n_patches.argtypes = [VP]
        L.dh_scene_n_patches.restype = I64
        L.dh_scene_culled.argtypes = [VP]
        L.dh_scene_culled.restype = U64
        L.dh_scene_tables.argtypes = [VP] * 6
        L.dh_scene_material.argtypes = [VP, C.c_int, VP, VP]
        L.dh_analytic_nearest.argtypes = [VP, I64, VP, VP, VP, VP]
        L.dh_newton_cull.argtypes = [VP, I64, VP, VP, VP, VP]
        L.dh_eye_pass_kernel.argtypes = [VP, C.POINTER(abi.SppmParams), C.POINTER(abi.SppmCameraDesc), U64, C.c_int, C.c_int,
                                         I64, VP, VP]
        L.dh_eye_pass_kernel.restype = I64
        L.dh_photon_trace_kernel.argtypes = [VP, C.POINTER(abi.SppmParams), U64, U64, U64, C.c_int, I64, VP, VP]
        L.dh_photon_trace_kernel.restype = I64
        L.dh_photon_trace_wavefront.argtypes = [VP, C.POINTER(abi.SppmParams), U64, U64, U64, C.c_int, C.c_int, I64, VP, VP, VP, I64, C.c_int]
        L.dh_photon_trace_wavefront.restype = I64
        L.dh_gather_filtered.argtypes = [I64, VP, C.POINTER(Grid), I64, VP, DBL, C.c_int, C.c_int, C.c_int, VP, VP, VP, VP, VP]
        L.dh_set_one_warp_blocks.argtypes = [C.c_int]
        L.dh_work_read.argtypes = [VP, C.c_int]
        L.dh_bounds.argtypes = [I64, VP, C.c_int, VP]
        L.dh_tonemap.argtypes = [VP, C.c_int, C.c_int, DBL, VP]
        L.dh_tex_expand.argtypes = [VP, C.c_int, C.c_int, VP, VP]
        L.dh_shared_pass.argtypes = [I64, C.c_int, DBL, DBL, DBL, VP, VP, VP, I64, VP, VP, VP, VP, VP, VP]
        L.dh_update_resolve.argtypes = [I64, C.c_int, DBL, C.c_int, VP, VP, VP, VP, VP, VP, VP, VP, C.c_int, C.c_int, C.c_int, VP, VP, VP]
        L.dh_intersect_kernel.argtypes = [VP, C.POINTER(abi.SppmParams), C.c_int, I64, VP, VP, U64, C.c_int, VP, VP]
        L.dh_patch_batch_kernel.argtypes = [VP, C.POINTER(abi.SppmParams), C.c_int, I64, VP, VP, VP, C.c_int, VP]
        assert L.dh_sizeof_hpraw() == hpraw_dtype.itemsize and L.dh_sizeof_deposit() == deposit_dtype.itemsize

    def draw2(self, seed, pass_id, ident, node, purpose, sub, patch):
        o = (C.c_double * 2)()
        self.lib.dh_draw2(seed, pass_id, ident, node, purpose, sub, patch, o)
        return o[0], o[1]

    def aabb_intersect(self, box6, rays6):
        box6, rays6 = _f64(box6, (6,)), _f64(rays6, (-1, 6))
        out = np.empty(rays6.shape[0], dtype=np.int32)
        self.lib.dh_aabb_intersect(_p(box6), rays6.shape[0], _p(rays6), _p(out))
        return out.astype(bool)

    def bezier_eval(self, mode, P48, uv):
        """mode 0: eval_exact(P), 1: eval_mono(A), 2: eval_fast(P) -> (S, Su, Sv) [n, 3] each."""
        P48, uv = _f64(P48, (48,)), _f64(uv, (-1, 2))
        out = np.empty((uv.shape[0], 9))
        self.lib.dh_bezier_eval(mode, _p(P48), uv.shape[0], _p(uv), _p(out))
        return out[:, 0:3], out[:, 3:6], out[:, 6:9]

    def newton(self, mode, P48, rays6, seeds, iters=10):
        """Bezier3::newton of the kernels on ONE patch (P for modes 0/2, the monomial A for mode 1)."""
        P48, rays6 = _f64(P48, (48,)), _f64(rays6, (-1, 6))
        n = rays6.shape[0]
        seeds = _f64(seeds, (n, -1, 2))
        out = np.empty((n, 3))
        self.lib.dh_newton(mode, _p(P48), n, _p(rays6), _p(seeds), seeds.shape[1], iters, _p(out))
        return out

    def color_uv(self, tex_map, uv, lod=0):
        tex_map = np.ascontiguousarray(tex_map, dtype=np.uint8)
        uv = _f64(uv, (-1, 2))
        out = np.empty((uv.shape[0], 3))
        self.lib.dh_color_uv(tex_map.shape[0], tex_map.shape[1], _p(tex_map), uv.shape[0], _p(uv), _p(out), lod)
        return out

    def unit_rand_refl(self, norm3, u12, fast=False):
        norm3, u12 = _f64(norm3, (-1, 3)), _f64(u12, (-1, 2))
        out = np.empty_like(norm3)
        self.lib.dh_unit_rand_refl(norm3.shape[0], _p(norm3), _p(u12), int(fast), _p(out))
        return out

    def refracted(self, I3, N3, nn):
        I3, N3, nn = _f64(I3, (-1, 3)), _f64(N3, (-1, 3)), _f64(nn, (-1,))
        out = np.empty_like(I3)
        self.lib.dh_refracted(I3.shape[0], _p(I3), _p(N3), _p(nn), _p(out))
        return out

    def unit_rand_vec(self, seed, pass_id, ids, purpose):
        ids = np.ascontiguousarray(ids, dtype=np.uint64)
        out = np.empty((ids.shape[0], 3))
        self.lib.dh_unit_rand_vec(seed, pass_id, ids.shape[0], _p(ids), purpose, _p(out))
        return out

    def gather_filtered(self, hp, dep, search_r2, count_pairs=True, blocks=16, exact=False, shuffle_seed=None):
        """build_index + deposit binning + gather_filtered_kernel (the product's kernels; CUB's sorts and
        scans restated with std:: algorithms) -> newN [n], tao [n, 3] in the order of `hp`, info, pairs."""
        n, m = hp["pos"].shape[0], dep["x"].shape[0]
        raw = np.zeros(n, dtype=hpraw_dtype)
        for k in ("pos", "n", "wgt", "r2"):
            raw[k] = hp[k]
        raw["okey"] = np.arange(n, dtype=np.uint64)
        if shuffle_seed is not None:     # arrival order of the eye pass: the canonical order is restored from okey
            perm = np.random.RandomState(shuffle_seed).permutation(n)
            raw = np.ascontiguousarray(raw[perm])
        d = np.zeros(max(m, 1), dtype=deposit_dtype)
        for k in ("x", "n", "flux"):
            d[k][:m] = dep[k]
        g = make_grid(_f64(hp["pos"]), search_r2)
        order = np.zeros(n, dtype=np.uint32)
        dN = np.zeros(n, dtype=np.int32)
        dtao = np.zeros((3, n))
        info = np.zeros(5, dtype=np.int64)
        pairs = np.zeros(2, dtype=np.uint64)
        rc = self.lib.dh_gather_filtered(n, _p(raw), C.byref(g), m, _p(d), search_r2, int(count_pairs), int(exact), blocks, _p(order),
                                         _p(dN), _p(dtao), _p(info), _p(pairs))
        assert rc == 0
        newN = np.zeros(n, dtype=np.int32)
        tao = np.zeros((n, 3))
        if shuffle_seed is not None:
            order = perm[order]
        newN[order] = dN
        tao[order] = dtao.T
        return newN, tao, dict(zip(("n_occ", "n_slots", "n_units", "n_items", "kept"), info.tolist())), pairs.tolist()

    def update_resolve(self, mode, alpha, N, newN, dN, r2, tao, dtao, pix, cnt, onetime, w, h, rounds=1, canonical="pixel"):
        """update_kernel x rounds, then hp_rank_kernel + hp_canon_kernel + resolve_kernel; tao / dtao [n, 3].  The
        canonical order is (pixel, index), the map's own order is the index.  In place on N, newN, r2; returns
        (tao [n, 3], film [h, w, 3])."""
        n = N.shape[0]
        t, dt = np.ascontiguousarray(tao.T), np.ascontiguousarray(dtao.T)
        row, col = np.ascontiguousarray(pix[:, 0], dtype=np.int32), np.ascontiguousarray(pix[:, 1], dtype=np.int32)
        film = np.zeros((h, w, 3))
        # canonical = "index": an imported map whose order separates a pixel's hit points (several runs per pixel)
        raw_of_rank = (np.argsort(row.astype(np.int64) * w + col, kind="stable") if canonical == "pixel" else np.arange(n)).astype(np.uint32)
        raw_of_pos = np.arange(n, dtype=np.uint32)
        self.lib.dh_update_resolve(n, mode, alpha, rounds, _p(N), _p(newN), _p(dN), _p(r2), _p(t), _p(dt), _p(row), _p(col),
                                   cnt, onetime, w, _p(film), _p(raw_of_rank), _p(raw_of_pos))
        return np.ascontiguousarray(t.T), film

    def shared_pass(self, r2, n, tau, pix, newN, tao, alpha, photons, init_r2=0.0):
        """One eye pass of the shared per-pixel statistics through shared_{init,load,accum,update,resolve}_kernel.
        r2, n [h, w] and tau [h, w, 3] in place; returns (radius each hit point measured with, film [h, w, 3])."""
        h, w = r2.shape
        t3 = np.ascontiguousarray(tau.reshape(h * w, 3).T)
        hp_r2 = np.zeros(newN.shape[0])
        film = np.zeros((h, w, 3))
        row, col = np.ascontiguousarray(pix[:, 0], dtype=np.int32), np.ascontiguousarray(pix[:, 1], dtype=np.int32)
        nn, tt = np.ascontiguousarray(newN, dtype=np.int32), np.ascontiguousarray(tao.T)
        self.lib.dh_shared_pass(h * w, w, init_r2, alpha, photons, _p(r2), _p(n), _p(t3), newN.shape[0], _p(row), _p(col),
                                _p(nn), _p(tt), _p(hp_r2), _p(film))
        tau[...] = t3.T.reshape(h, w, 3)
        return hp_r2, film

    def tonemap(self, film, div):
        film = _f64(film)
        h, w = film.shape[:2]
        out = np.zeros((h, w, 3), dtype=np.uint8)
        self.lib.dh_tonemap(_p(film), w, h, div, _p(out))
        return out

    def tex_expand(self, bgr8):
        """tex_expand_kernel + tex_mip_kernel as sppm_scene_upload runs them -> the levels of the chain, untiled:
        a list of uint32 arrays [rows_l, cols_l] (B | G << 8 | R << 16 per texel)."""
        bgr8 = np.ascontiguousarray(bgr8, dtype=np.uint8)
        rows, cols = bgr8.shape[:2]
        out = np.zeros(2 * rows * cols + 64, dtype=np.uint32)
        dims = np.zeros((8, 2), dtype=np.int32)
        n = self.lib.dh_tex_expand(_p(bgr8), rows, cols, _p(out), _p(dims))
        levels, o = [], 0
        for l in range(n):
            r, c = int(dims[l, 0]), int(dims[l, 1])
            levels.append(out[o:o + r * c].reshape(r, c).copy())
            o += r * c
        return levels

    def bounds(self, pos, blocks=7):
        raw = np.zeros(pos.shape[0], dtype=hpraw_dtype)
        raw["pos"] = pos
        out = np.zeros(9)
        self.lib.dh_bounds(pos.shape[0], _p(raw), blocks, _p(out))
        return out[:3], out[3:6], out[6:]

    WORK = ("items", "slices", "visits", "kept", "pairs", "pair_slots", "redo_lanes", "full_accepts")

    def gather_work(self, reset=True):
        """Algorithmic work counters of gather_filtered_kernel since the last reset (emulation only)."""
        out = np.zeros(len(self.WORK), dtype=np.uint64)
        self.lib.dh_work_read(_p(out), int(reset))
        return dict(zip(self.WORK, out.tolist()))

    def one_warp_blocks(self, on: bool):
        """Launch with one-warp blocks instead of the product's block sizes (a second schedule)."""
        self.lib.dh_set_one_warp_blocks(int(on))

    def scene(self, scene) -> HostScene:
        return HostScene(self, scene)
