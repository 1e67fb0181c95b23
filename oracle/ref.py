"""ctypes wrapper of oracle/_ref/libref.so — the reference's OWN code (built by
oracle/ref_shim/build_ref.py from /root/reference/raytracer/*.h against a thin OpenCV/FLANN
shim).  TEST INFRASTRUCTURE: pins the oracle (tests/test_ref_pin_cpu.py) and times the
reference's own PhotonMap for bench.py's cpu_baseline (time_photon_map)."""
from __future__ import annotations

import ctypes as C
import os
import tempfile

import numpy as np

from raytrace3_b200 import abi
from raytrace3_b200.api import hit_dtype
from raytrace3_b200.scene import Scene

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libref.so")
VP, I32, I64, DBL = C.c_void_p, C.c_int32, C.c_int64, C.c_double


def available() -> bool:
    return os.path.exists(LIB_PATH)


def _p(a):
    return None if a is None else a.ctypes.data


def _f64(a, shape=None):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a.reshape(shape) if shape is not None else a


class Reference:
    def __init__(self):
        self.lib = L = C.CDLL(LIB_PATH)
        L.ref_push_uniforms.argtypes = [VP, C.c_int]
        L.ref_underflows.restype = C.c_long
        L.ref_uniforms_left.restype = C.c_long
        L.ref_aabb_intersect.argtypes = [VP, VP, VP]
        L.ref_bezier_eval.argtypes = [VP, DBL, DBL, VP, VP, VP]
        L.ref_newton.argtypes = [VP, VP, VP, VP]
        L.ref_unit_rand_refl.argtypes = [VP, VP]
        L.ref_unit_rand_vec.argtypes = [VP]
        L.ref_refracted.argtypes = [VP, VP, DBL, VP]
        L.ref_camera.argtypes = [VP, VP, C.c_int, C.c_int, C.POINTER(abi.SppmCameraDesc), C.c_int, C.c_int, VP]
        L.ref_to_int.argtypes = [DBL]
        L.ref_register_image.argtypes = [C.c_char_p, C.c_int, C.c_int, VP]
        L.ref_color_uv.argtypes = [C.c_char_p, DBL, DBL, VP]
        L.ref_scene_create.argtypes = [C.POINTER(abi.SppmSceneDesc), C.c_char_p]
        L.ref_scene_create.restype = VP
        L.ref_scene_destroy.argtypes = [VP]
        L.ref_near_collide.argtypes = [VP, VP, VP]
        L.ref_photon.argtypes = [VP, C.c_int, I64, VP, VP, VP, VP, VP, VP]
        L.ref_view_trace.argtypes = [VP, VP, VP, C.c_int, C.c_int, C.c_int, C.c_int, I64, VP, VP, VP, VP, VP]
        L.ref_view_trace.restype = I64
        L.ref_update.argtypes = [I64, VP, VP, VP, VP]
        L.ref_resolve_one.argtypes = [VP, DBL, C.c_int, C.c_int, VP]
        self.can_time = hasattr(L, "ref_time_photon_map")  # a libref.so built before the timing entry existed
        if self.can_time:
            L.ref_flann_selfcheck.argtypes = [C.c_int, C.c_int, DBL, DBL, C.c_uint, VP]
            L.ref_flann_selfcheck.restype = C.c_long
            L.ref_time_photon_map.argtypes = [VP, VP, VP, C.c_int, C.c_int, C.c_int, C.c_int, DBL, VP]
        self.u_first = bool(L.ref_arg_order_u_first())
        self._keep = []
        self._tmp = tempfile.TemporaryDirectory()

    # -- injected random numbers
    def feed(self, vals, kinds=None):
        """Queue uniforms in the reference's consumption order.  Newton (u0, v0) pairs
        (kinds 1, 2) are swapped when this build evaluates the second GetRandu first (Q16)."""
        vals = np.array(vals, dtype=np.float64)
        if kinds is not None and not self.u_first:
            k = np.asarray(kinds)
            i = np.nonzero(k == 1)[0]
            assert (k[i + 1] == 2).all()
            vals[i], vals[i + 1] = vals[i + 1].copy(), vals[i].copy()
        self.lib.ref_clear_uniforms()
        self.lib.ref_push_uniforms(_p(vals), len(vals))

    def check_consumed(self):
        assert self.lib.ref_underflows() == 0, "the reference asked for more random numbers than were fed"
        assert self.lib.ref_uniforms_left() == 0, "the reference consumed fewer random numbers than were fed"

    # -- pieces
    def aabb_intersect(self, box6, org, dirv):
        b, o, d = _f64(box6, (6,)), _f64(org, (3,)), _f64(dirv, (3,))
        return bool(self.lib.ref_aabb_intersect(_p(b), _p(o), _p(d)))

    def bezier_eval(self, P, u, v):
        P = _f64(P, (48,))
        S, Su, Sv = np.empty(3), np.empty(3), np.empty(3)
        self.lib.ref_bezier_eval(_p(P), u, v, _p(S), _p(Su), _p(Sv))
        return S, Su, Sv

    def newton(self, P, org, dirv, seeds_uv):
        s = _f64(seeds_uv, (-1, 2))
        assert s.shape[0] == 50
        self.feed(s.reshape(-1), np.tile([1, 2], 50))
        P, o, d = _f64(P, (48,)), _f64(org, (3,)), _f64(dirv, (3,))
        out = np.empty(3)
        self.lib.ref_newton(_p(P), _p(o), _p(d), _p(out))
        self.check_consumed()
        return out

    def unit_rand_refl(self, norm, u1, u2):
        self.feed([u1, u2])
        n, o = _f64(norm, (3,)), np.empty(3)
        self.lib.ref_unit_rand_refl(_p(n), _p(o))
        self.check_consumed()
        return o

    def refracted(self, I, N, n):
        I, N, o = _f64(I, (3,)), _f64(N, (3,)), np.empty(3)
        self.lib.ref_refracted(_p(I), _p(N), n, _p(o))
        return o

    def camera(self, pos, look, w, h, y=0, x=0):
        pos, look = _f64(pos, (3,)), _f64(look, (3,))
        cam = abi.SppmCameraDesc()
        ray = np.empty(6)
        self.lib.ref_camera(_p(pos), _p(look), w, h, C.byref(cam), y, x, _p(ray))
        return cam, ray

    def to_int(self, x):
        return self.lib.ref_to_int(float(x))

    def color_uv(self, tex_map, u, v):
        name = f"img{len(self._keep)}".encode()
        m = np.ascontiguousarray(tex_map, dtype=np.uint8)
        self._keep.append(m)
        self.lib.ref_register_image(name, m.shape[0], m.shape[1], _p(m))
        o = np.empty(3)
        self.lib.ref_color_uv(name, u, v, _p(o))
        return o

    # -- scenes built from the reference's own classes
    def scene(self, scene):
        d = scene.flatten() if isinstance(scene, Scene) else scene
        self._keep.append((scene, d))
        h = self.lib.ref_scene_create(C.byref(d), self._tmp.name.encode())
        return h

    def near_collide(self, h, ray6, vals=(), kinds=None):
        self.feed(vals, kinds)
        r = _f64(ray6, (6,))
        out = np.zeros(1, dtype=hit_dtype)
        self.lib.ref_near_collide(h, _p(r), _p(out))
        self.check_consumed()
        return out[0]

    def photon(self, h, n_lights, hp, vals, kinds):
        """One photon from each light in turn (Raytracer.h:226-232), accumulating newN / tao."""
        self.feed(vals, kinds)
        a = [_f64(hp[k]) for k in ("pos", "n", "wgt", "r2")]
        n = a[0].shape[0]
        newN, tao = np.zeros(n, dtype=np.int32), np.zeros((n, 3))
        for light in range(n_lights):
            self.lib.ref_photon(h, light, n, _p(a[0]), _p(a[1]), _p(a[2]), _p(a[3]), _p(newN), _p(tao))
        self.check_consumed()
        return newN, tao

    def view_trace(self, h, cam_pos, cam_look, w, hgt, i, j, vals=(), kinds=None, cap=64):
        self.feed(vals, kinds)
        pos, look = _f64(cam_pos, (3,)), _f64(cam_look, (3,))
        p3, n3, w3 = np.zeros((cap, 3)), np.zeros((cap, 3)), np.zeros((cap, 3))
        r2 = np.zeros(cap)
        pix = np.zeros((cap, 2), dtype=np.int32)
        n = self.lib.ref_view_trace(h, _p(pos), _p(look), w, hgt, i, j, cap, _p(p3), _p(n3), _p(w3), _p(r2), _p(pix))
        self.check_consumed()
        return dict(pos=p3[:n], n=n3[:n], wgt=w3[:n], r2=r2[:n], pix=pix[:n])

    def update(self, newN, N, r2, tao):
        self.lib.ref_update(newN.shape[0], _p(newN), _p(N), _p(r2), _p(tao))

    def resolve_one(self, tao, r2, cnt, onetime):
        t, o = _f64(tao, (3,)), np.empty(3)
        self.lib.ref_resolve_one(_p(t), float(r2), cnt, onetime, _p(o))
        return o

    # -- the shim's kd-tree against its exhaustive scan; returns (queries that differ, neighbours compared)
    def flann_selfcheck(self, n, nq, side, radius2, seed=1):
        found = C.c_long(0)
        bad = self.lib.ref_flann_selfcheck(n, nq, float(side), float(radius2), seed, C.addressof(found))
        return int(bad), int(found.value)

    # -- the reference's own GetHitPoint + index build (untimed) and PhotonMap(cnt) (timed), one replica per thread
    def time_photon_map(self, h, cam_pos, cam_look, w, hgt, cnt, threads, target_s=0.0):
        """target_s > 0: the round count is chosen from a 10-round probe (cnt is the upper limit)."""
        pos, look, out = _f64(cam_pos, (3,)), _f64(cam_look, (3,)), np.zeros(5)
        self.lib.ref_time_photon_map(h, _p(pos), _p(look), w, hgt, cnt, threads, float(target_s), _p(out))
        return dict(setup_s=float(out[0]), map_s=float(out[1]), hit_points=int(out[2]), photons=int(out[3]), rounds=int(out[4]))
