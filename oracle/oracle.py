"""ctypes wrapper of oracle/_build/liboracle.so (see sppm_oracle.h).  TEST INFRASTRUCTURE."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from raytrace3_b200 import abi
from raytrace3_b200.api import hit_dtype
from raytrace3_b200.scene import Camera, Scene

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_build", "liboracle.so")

VP, I32, I64, U64, U32, DBL = C.c_void_p, C.c_int32, C.c_int64, C.c_uint64, C.c_uint32, C.c_double


def build_oracle(force: bool = False) -> str:
    srcs = [os.path.join(_HERE, f) for f in ("sppm_oracle.cpp", "sppm_oracle.h")] + [
        os.path.join(_HERE, "..", "include", "sppm_b200.h")]
    stale = force or not os.path.exists(LIB_PATH) or any(
        os.path.getmtime(s) > os.path.getmtime(LIB_PATH) for s in srcs)
    if stale:
        r = subprocess.run(["make", "-C", _HERE] + (["-B"] if force else []), capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("oracle build failed:\n" + r.stdout + r.stderr)
    return LIB_PATH


def _p(a):
    return None if a is None else a.ctypes.data


def _f64(a, shape=None):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a.reshape(shape) if shape is not None else a


class Oracle:
    def __init__(self):
        # SPPM_ORACLE_LIB: an instrumented build of the same source (`make -C oracle asan`)
        self.lib = C.CDLL(os.environ.get("SPPM_ORACLE_LIB") or build_oracle())
        L = self.lib
        L.orc_draw2.argtypes = [U64, U64, U64, U32, U32, U32, U32, VP]
        L.orc_newton.argtypes = [VP, VP, VP, VP, C.c_int, C.c_int, VP]
        L.orc_bezier_eval.argtypes = [VP, DBL, DBL, VP, VP, VP]
        L.orc_color_uv.argtypes = [C.POINTER(abi.SppmTexture), DBL, DBL, VP]
        L.orc_unit_rand_refl.argtypes = [VP, DBL, DBL, VP]
        L.orc_refracted.argtypes = [VP, VP, DBL, VP]
        L.orc_camera_lookat.argtypes = [VP, VP, C.c_int, C.c_int, C.POINTER(abi.SppmCameraDesc)]
        L.orc_camera_emit.argtypes = [C.POINTER(abi.SppmCameraDesc), C.c_int, C.c_int, VP]
        L.orc_intersect_batch.argtypes = [C.POINTER(abi.SppmSceneDesc), C.POINTER(abi.SppmParams), VP, I64, VP, U64, I32, I32, VP]
        L.orc_bezier_patch_batch.argtypes = [C.POINTER(abi.SppmSceneDesc), C.POINTER(abi.SppmParams), VP, VP, I64, VP, VP]
        L.orc_eye_pass.argtypes = [C.POINTER(abi.SppmSceneDesc), C.POINTER(abi.SppmParams), C.POINTER(abi.SppmCameraDesc),
                                   U64, I64, VP, VP, VP, VP, VP, VP]
        L.orc_eye_pass.restype = I64
        L.orc_view_trace_pixel.argtypes = [C.POINTER(abi.SppmSceneDesc), C.POINTER(abi.SppmParams), C.POINTER(abi.SppmCameraDesc),
                                           U64, C.c_int, C.c_int, I64, VP, VP, VP, VP, VP]
        L.orc_view_trace_pixel.restype = I64
        L.orc_gather_brute.argtypes = [I64, VP, VP, VP, VP, I64, VP, VP, VP, DBL, VP, VP, VP, VP]
        L.orc_kdtree_build.argtypes = [I64, VP]
        L.orc_kdtree_build.restype = VP
        L.orc_kdtree_free.argtypes = [VP]
        L.orc_gather_kdtree.argtypes = [VP, I64, VP, VP, VP, VP, I64, VP, VP, VP, DBL, VP, VP]
        L.orc_photon_pass.argtypes = [C.POINTER(abi.SppmSceneDesc), C.POINTER(abi.SppmParams), U64, U64, U64, VP, I64,
                                      VP, VP, VP, VP, VP, VP, I64, VP, VP, VP, VP, VP, VP]
        L.orc_photon_pass.restype = I64
        L.orc_photon_bench.argtypes = [C.POINTER(abi.SppmSceneDesc), C.POINTER(abi.SppmParams), U64, U64, I32, VP, I64,
                                       VP, VP, VP, VP, VP]
        L.orc_photon_bench.restype = DBL
        L.orc_patch_bench.argtypes = [C.POINTER(abi.SppmSceneDesc), C.POINTER(abi.SppmParams), VP, VP, I64, I32, VP]
        L.orc_patch_bench.restype = DBL
        L.orc_update.argtypes = [I64, I32, DBL, VP, VP, VP, VP]
        L.orc_resolve.argtypes = [I64, VP, VP, VP, I32, I32, I32, I32, VP]
        L.orc_tonemap.argtypes = [VP, I32, I32, DBL, VP]
        L.orc_render_replicas.argtypes = [C.POINTER(abi.SppmSceneDesc), C.POINTER(abi.SppmParams), VP, VP, I32, I32, I32,
                                          I32, I32, DBL, VP, VP, VP]
        L.orc_max_threads.restype = C.c_int
        L.orc_log_end.argtypes = [VP, VP, I64]
        L.orc_log_end.restype = I64

    # -- small pieces
    def philox(self, ctr, key):
        c = (C.c_uint32 * 4)(*ctr)
        k = (C.c_uint32 * 2)(*key)
        o = (C.c_uint32 * 4)()
        self.lib.orc_philox4x32_10(c, k, o)
        return tuple(o)

    def draw2(self, seed, pass_id, ident, node, purpose, sub, patch):
        o = (C.c_double * 2)()
        self.lib.orc_draw2(seed, pass_id, ident, node, purpose, sub, patch, o)
        return o[0], o[1]

    def inv3(self, m):
        m = _f64(m, (9,))
        o = np.empty(9)
        self.lib.orc_inv3(C.c_void_p(m.ctypes.data), C.c_void_p(o.ctypes.data))
        return o.reshape(3, 3)

    def aabb_intersect(self, box6, org, dirv):
        b, o, d = _f64(box6, (6,)), _f64(org, (3,)), _f64(dirv, (3,))
        return bool(self.lib.orc_aabb_intersect(C.c_void_p(b.ctypes.data), C.c_void_p(o.ctypes.data), C.c_void_p(d.ctypes.data)))

    def bezier_eval(self, P, u, v):
        P = _f64(P, (48,))
        S, Su, Sv = np.empty(3), np.empty(3), np.empty(3)
        self.lib.orc_bezier_eval(_p(P), u, v, _p(S), _p(Su), _p(Sv))
        return S, Su, Sv

    def newton(self, P, org, dirv, seeds_uv, iters=10):
        P, o, d = _f64(P, (48,)), _f64(org, (3,)), _f64(dirv, (3,))
        s = _f64(seeds_uv, (-1, 2))
        out = np.empty(3)
        self.lib.orc_newton(_p(P), _p(o), _p(d), _p(s), s.shape[0], iters, _p(out))
        return out

    def color_uv(self, tex_map, u, v):
        t = abi.SppmTexture(rows=tex_map.shape[0], cols=tex_map.shape[1], bgr8=tex_map.ctypes.data)
        o = np.empty(3)
        self.lib.orc_color_uv(C.byref(t), u, v, _p(o))
        return o

    def unit_rand_refl(self, norm, u1, u2):
        n, o = _f64(norm, (3,)), np.empty(3)
        self.lib.orc_unit_rand_refl(_p(n), u1, u2, _p(o))
        return o

    def refracted(self, I, N, n):
        I, N, o = _f64(I, (3,)), _f64(N, (3,)), np.empty(3)
        self.lib.orc_refracted(_p(I), _p(N), n, _p(o))
        return o

    def camera_lookat(self, pos, look, w, h) -> abi.SppmCameraDesc:
        pos, look = _f64(pos, (3,)), _f64(look, (3,))
        cam = abi.SppmCameraDesc()
        self.lib.orc_camera_lookat(_p(pos), _p(look), w, h, C.byref(cam))
        return cam

    def camera_emit(self, cam, y, x):
        o = np.empty(6)
        self.lib.orc_camera_emit(C.byref(cam), y, x, _p(o))
        return o

    # -- batches
    @staticmethod
    def _desc(scene):
        return scene.flatten() if isinstance(scene, Scene) else scene

    def intersect_batch(self, scene, prm, rays6, ids=None, pass_id=0, node=0, domain=abi.SPPM_DOMAIN_TEST):
        d = self._desc(scene)
        rays6 = _f64(rays6, (-1, 6))
        out = np.zeros(rays6.shape[0], dtype=hit_dtype)
        if ids is not None:
            ids = np.ascontiguousarray(ids, dtype=np.uint64)
        self.lib.orc_intersect_batch(C.byref(d), C.byref(prm), _p(rays6), rays6.shape[0], _p(ids), pass_id, node, domain, _p(out))
        return out

    def bezier_patch_batch(self, scene, prm, rays6, patch_idx, seeds_uv=None):
        d = self._desc(scene)
        rays6 = _f64(rays6, (-1, 6))
        n = rays6.shape[0]
        patch_idx = np.ascontiguousarray(patch_idx, dtype=np.int32)
        if seeds_uv is not None:
            seeds_uv = _f64(seeds_uv, (n, prm.newton_restarts, 2))
        out = np.empty((n, 3))
        self.lib.orc_bezier_patch_batch(C.byref(d), C.byref(prm), _p(rays6), _p(patch_idx), n, _p(seeds_uv), _p(out))
        return out

    def eye_pass(self, scene, prm, cam, pass_id=0, cap=None):
        d = self._desc(scene)
        cd = cam.desc() if isinstance(cam, Camera) else cam
        if cap is None:
            cap = cd.w * cd.h * 4 + 1024
        pos3, n3, wgt3 = (np.zeros((cap, 3)) for _ in range(3))
        r2 = np.zeros(cap)
        pix = np.zeros((cap, 2), dtype=np.int32)
        node = np.zeros(cap, dtype=np.uint32)
        n = self.lib.orc_eye_pass(C.byref(d), C.byref(prm), C.byref(cd), pass_id, cap, _p(pos3), _p(n3), _p(wgt3), _p(r2),
                                  _p(pix), _p(node))
        if n > cap:   # branching materials (Raytracer.h:312-335): the count is returned, trace again with room for it
            return self.eye_pass(scene, prm, cam, pass_id, cap=n)
        return dict(pos=pos3[:n].copy(), n=n3[:n].copy(), wgt=wgt3[:n].copy(), r2=r2[:n].copy(), pix=pix[:n].copy(),
                    node=node[:n].copy())

    def view_trace_pixel(self, scene, prm, cam, i, j, pass_id=0, cap=64):
        d = self._desc(scene)
        cd = cam.desc() if isinstance(cam, Camera) else cam
        p3, n3, w3 = np.zeros((cap, 3)), np.zeros((cap, 3)), np.zeros((cap, 3))
        r2 = np.zeros(cap)
        pix = np.zeros((cap, 2), dtype=np.int32)
        n = self.lib.orc_view_trace_pixel(C.byref(d), C.byref(prm), C.byref(cd), pass_id, i, j, cap, _p(p3), _p(n3), _p(w3),
                                          _p(r2), _p(pix))
        return dict(pos=p3[:n], n=n3[:n], wgt=w3[:n], r2=r2[:n], pix=pix[:n])

    def gather_brute(self, hp, dep, search_r2, newN=None, tao=None):
        n = hp["pos"].shape[0]
        m = dep["x"].shape[0]
        newN = np.zeros(n, dtype=np.int32) if newN is None else newN
        tao = np.zeros((n, 3)) if tao is None else tao
        cand, acc = C.c_uint64(), C.c_uint64()
        a = [_f64(hp[k]) for k in ("pos", "n", "wgt", "r2")]
        b = [_f64(dep[k]) for k in ("x", "n", "flux")]
        self.lib.orc_gather_brute(n, _p(a[0]), _p(a[1]), _p(a[2]), _p(a[3]), m, _p(b[0]), _p(b[1]), _p(b[2]), search_r2,
                                  _p(newN), _p(tao), C.addressof(cand), C.addressof(acc))
        return newN, tao, cand.value, acc.value

    def gather_kdtree(self, hp, dep, search_r2):
        n = hp["pos"].shape[0]
        m = dep["x"].shape[0]
        newN, tao = np.zeros(n, dtype=np.int32), np.zeros((n, 3))
        a = [_f64(hp[k]) for k in ("pos", "n", "wgt", "r2")]
        b = [_f64(dep[k]) for k in ("x", "n", "flux")]
        tree = self.lib.orc_kdtree_build(n, _p(a[0]))
        try:
            self.lib.orc_gather_kdtree(tree, n, _p(a[0]), _p(a[1]), _p(a[2]), _p(a[3]), m, _p(b[0]), _p(b[1]), _p(b[2]),
                                       search_r2, _p(newN), _p(tao))
        finally:
            self.lib.orc_kdtree_free(tree)
        return newN, tao

    def photon_pass(self, scene, prm, begin, count, pass_id, hp=None, newN=None, tao=None, keep_deposits=False,
                    use_kdtree=True):
        d = self._desc(scene)
        n = 0 if hp is None else hp["pos"].shape[0]
        a = [None] * 4 if hp is None else [_f64(hp[k]) for k in ("pos", "n", "wgt", "r2")]
        newN = np.zeros(max(n, 1), dtype=np.int32) if newN is None else newN
        tao = np.zeros((max(n, 1), 3)) if tao is None else tao
        tree = self.lib.orc_kdtree_build(n, _p(a[0])) if (n and use_kdtree) else None
        cap = int(count * max(1, d.n_lights) * (prm.max_dep + 1)) if keep_deposits else 0
        dx, dn, df = (np.zeros((cap, 3)) for _ in range(3)) if keep_deposits else (None, None, None)
        dpid = np.zeros(cap, dtype=np.uint64) if keep_deposits else None
        dbn = np.zeros(cap, dtype=np.int32) if keep_deposits else None
        counters = np.zeros(4, dtype=np.uint64)
        try:
            m = self.lib.orc_photon_pass(C.byref(d), C.byref(prm), begin, count, pass_id, tree, n, _p(a[0]), _p(a[1]),
                                         _p(a[2]), _p(a[3]), _p(newN), _p(tao), cap, _p(dx), _p(dn), _p(df), _p(dpid),
                                         _p(dbn), _p(counters))
        finally:
            if tree:
                self.lib.orc_kdtree_free(tree)
        out = dict(newN=newN[:n], tao=tao[:n], counters=dict(zip(("queries", "solves", "deposits", "photons"), counters.tolist())))
        if keep_deposits:
            out["dep"] = dict(x=dx[:m], n=dn[:m], flux=df[:m], photon=dpid[:m], bounce=dbn[:m])
        return out

    def photon_bench(self, scene, prm, hp, count_per_thread, threads, pass_id=0):
        """Wall seconds for `threads` replicas x count_per_thread photons (per light) + counters."""
        d = self._desc(scene)
        a = [_f64(hp[k]) for k in ("pos", "n", "wgt", "r2")]
        n = a[0].shape[0]
        tree = self.lib.orc_kdtree_build(n, _p(a[0]))
        counters = np.zeros(4, dtype=np.uint64)
        try:
            s = self.lib.orc_photon_bench(C.byref(d), C.byref(prm), count_per_thread, pass_id, threads, tree, n, _p(a[0]),
                                          _p(a[1]), _p(a[2]), _p(a[3]), _p(counters))
        finally:
            self.lib.orc_kdtree_free(tree)
        return s, dict(zip(("queries", "solves", "deposits", "photons"), counters.tolist()))

    def patch_bench(self, scene, prm, rays6, patch_idx, threads):
        d = self._desc(scene)
        rays6 = _f64(rays6, (-1, 6))
        patch_idx = np.ascontiguousarray(patch_idx, dtype=np.int32)
        out = np.empty((rays6.shape[0], 3))
        s = self.lib.orc_patch_bench(C.byref(d), C.byref(prm), _p(rays6), _p(patch_idx), rays6.shape[0], threads, _p(out))
        return s, out

    def update(self, mode, alpha, newN, N, r2, tao):
        self.lib.orc_update(newN.shape[0], mode, alpha, _p(newN), _p(N), _p(r2), _p(tao))

    @staticmethod
    def shared_fold(r2, n, tau, pix, newN, tao, alpha):
        """Per-pixel SPPM statistics update (Hachisuka & Jensen 2009, eq. 9-12; the estimator the
        reference's README.md:9-17 names but Raytracer.h:69-79,429-458 does not implement):
        M = sum newN, Phi = sum tao over the pixel's hit points; N' = N + alpha M;
        R2' = R2 (N + alpha M)/(N + M); tau' = (tau + Phi) R2'/R2.  In place on r2, n [h,w], tau [h,w,3]."""
        h, w = r2.shape
        m = np.zeros((h, w), dtype=np.int64)
        phi = np.zeros((h, w, 3))
        np.add.at(m, (pix[:, 0], pix[:, 1]), newN)
        np.add.at(phi, (pix[:, 0], pix[:, 1]), tao)
        hit = m > 0
        n1 = n + alpha * m
        k = np.where(hit, n1 / np.where(hit, n + m, 1.0), 1.0)
        n[...] = np.where(hit, n1, n)
        r2 *= k
        tau[...] = np.where(hit[..., None], (tau + phi) * k[..., None], tau)
        return m, phi

    def resolve(self, pix, r2, tao, cnt, onetime, w, h):
        film = np.zeros((h, w, 3))
        pix = np.ascontiguousarray(pix, dtype=np.int32)
        r2, tao = _f64(r2), _f64(tao)
        self.lib.orc_resolve(pix.shape[0], _p(pix), _p(r2), _p(tao), cnt, onetime, w, h, _p(film))
        return film

    def tonemap(self, film, div):
        film = _f64(film)
        h, w = film.shape[:2]
        out = np.zeros((h, w, 3), dtype=np.uint8)
        self.lib.orc_tonemap(_p(film), w, h, div, _p(out))
        return out

    def log_begin(self):
        self.lib.orc_log_begin()

    def log_end(self, cap=1 << 20):
        vals = np.empty(cap)
        kinds = np.empty(cap, dtype=np.int8)
        n = self.lib.orc_log_end(_p(vals), _p(kinds), cap)
        assert n <= cap
        return vals[:n].copy(), kinds[:n].copy()

    def max_threads(self):
        return self.lib.orc_max_threads()

    def render_replicas(self, scene, prm, cam_pos, cam_look, w, h, cnt, onetime, threads, jitter=0.00015, want_film=True):
        d = self._desc(scene)
        pos, look = _f64(cam_pos, (3,)), _f64(cam_look, (3,))
        film = np.zeros((h, w, 3)) if want_film else None
        times = np.zeros(4)
        counters = np.zeros(4, dtype=np.uint64)
        self.lib.orc_render_replicas(C.byref(d), C.byref(prm), _p(pos), _p(look), w, h, cnt, onetime, threads, jitter,
                                     _p(film), _p(times), _p(counters))
        return dict(film=film, eye_s=times[0], build_s=times[1], photon_s=times[2], total_s=times[3],
                    counters=dict(zip(("queries", "solves", "deposits", "photons"), counters.tolist())))
