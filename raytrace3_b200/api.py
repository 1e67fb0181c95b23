"""Python host side above the C ABI: `Context` (thin RAII wrapper of sppm_ctx) and the
reference-facing drivers `RayTracer` / `SPPMRayTracer` (Raytracer.h:82-419, 421-477).

Everything that computes goes through libsppm_b200.so; a missing library or device raises.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import abi
from .scene import Camera, Scene


class SppmError(RuntimeError):
    pass


def _ptr(a):
    return None if a is None else a.ctypes.data


def _f64(a, shape=None):
    a = np.ascontiguousarray(a, dtype=np.float64)
    if shape is not None:
        a = a.reshape(shape)
    return a


def photon_range(onetime: int, rank: int, world: int):
    """Contiguous photon-index range [begin, begin+count) of `rank` out of `world`
    (photons of one round are split across GPUs; ids stay global)."""
    base, rem = divmod(int(onetime), int(world))
    begin = rank * base + min(rank, rem)
    return begin, base + (1 if rank < rem else 0)


FUSE_PHOTONS = 1 << 28


def rounds_per_call(onetime: int, fuse: bool) -> int:
    """Rounds of `onetime` photons handed to one photon pass by RayTracer.PhotonMap (twin of
    sppm_host::rounds_per_call): whole rounds, at most FUSE_PHOTONS photons per light and call,
    one round when the rounds are not independent."""
    if not fuse or onetime <= 0:
        return 1
    return max(1, FUSE_PHOTONS // int(onetime))


def photon_calls(cnt: int, onetime: int, fuse: bool, rank: int = 0, world: int = 1):
    """(begin, count) of every photon pass `rank` issues for PhotonMap(cnt): photon k of round i is
    photon i*onetime + k of the eye pass (ids stay global, any fusion / any world traces the same set)."""
    step = rounds_per_call(onetime, fuse)
    out = []
    for i in range(0, int(cnt), step):
        nr = min(step, cnt - i)
        b, c = photon_range(nr * onetime, rank, world)
        out.append((i * onetime + b, c))
    return out


def row_range(h: int, rank: int, world: int):
    """Film rows [begin, end) `rank` traces in a sharded eye pass: contiguous, disjoint,
    covering [0, h); feed to Context.eye_pass_rows, then Context.allgather_hitpoints."""
    return int(h) * rank // world, int(h) * (rank + 1) // world


def host_pin(a: np.ndarray) -> bool:
    """Page-lock a C-contiguous array the caller keeps alive (sppm_host_pin): the film read-back of `resolve(out=a)` and
    texture uploads then copy by DMA.  Best effort: False if the driver refuses; pair with host_unpin."""
    assert a.flags.c_contiguous
    return abi.load_library().sppm_host_pin(a.ctypes.data, a.nbytes) == 0


def host_unpin(a: np.ndarray) -> bool:
    return abi.load_library().sppm_host_unpin(a.ctypes.data) == 0


class Context:
    """One sppm_ctx = one CUDA device.  Methods map 1:1 onto include/sppm_b200.h."""

    def __init__(self, device: int = 0, params: abi.SppmParams | None = None):
        self.lib = abi.load_library()
        h = C.c_void_p()
        rc = self.lib.sppm_ctx_create(device, C.byref(h))
        if rc != 0:
            raise SppmError(f"sppm_ctx_create({device}) failed [{rc}]: {self.lib.sppm_last_error(None).decode()}")
        self.h = h
        self.params = abi.default_params()
        if params is not None:
            self.set_params(params)
        self._scene_keep = None

    def close(self):
        if getattr(self, "h", None):
            self.lib.sppm_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _ck(self, rc, what):
        if rc != 0:
            raise SppmError(f"{what} failed [{rc}]: {self.lib.sppm_last_error(self.h).decode()}")

    # -- configuration
    def set_params(self, p: abi.SppmParams):
        self._ck(self.lib.sppm_set_params(self.h, C.byref(p)), "sppm_set_params")
        self.params = p

    def upload_scene(self, scene: Scene | abi.SppmSceneDesc):
        desc = scene.flatten() if isinstance(scene, Scene) else scene
        self._scene_keep = (scene, desc)
        self._ck(self.lib.sppm_scene_upload(self.h, C.byref(desc)), "sppm_scene_upload")

    # -- render path
    def eye_pass(self, cam: Camera | abi.SppmCameraDesc, pass_id: int = 0) -> int:
        d = cam.desc() if isinstance(cam, Camera) else cam
        n = C.c_int64()
        self._ck(self.lib.sppm_eye_pass(self.h, C.byref(d), pass_id, C.byref(n)), "sppm_eye_pass")
        return n.value

    def eye_pass_rows(self, cam, pass_id: int, row_begin: int, row_end: int, append: bool = False) -> int:
        """Eye-pass sharding: only film rows [row_begin, row_end)."""
        desc = cam.desc() if isinstance(cam, Camera) else cam
        n = C.c_int64()
        self._ck(self.lib.sppm_eye_pass_rows(self.h, C.byref(desc), pass_id, row_begin, row_end, 1 if append else 0, C.byref(n)),
                 "sppm_eye_pass_rows")
        return n.value

    def allgather_hitpoints(self) -> int:
        n = C.c_int64()
        self._ck(self.lib.sppm_allgather_hitpoints(self.h, C.byref(n)), "sppm_allgather_hitpoints")
        return n.value

    def build_index(self):
        self._ck(self.lib.sppm_build_index(self.h), "sppm_build_index")

    def photon_pass(self, begin: int, count: int, pass_id: int):
        self._ck(self.lib.sppm_photon_pass(self.h, begin, count, pass_id), "sppm_photon_pass")

    def allreduce(self):
        self._ck(self.lib.sppm_allreduce(self.h), "sppm_allreduce")

    def update(self):
        self._ck(self.lib.sppm_update(self.h), "sppm_update")

    def resolve(self, cnt: int, onetime: int, w: int, h: int, out: np.ndarray | None = None) -> np.ndarray:
        film = out if out is not None else np.empty((h, w, 3), dtype=np.float64)
        assert film.flags.c_contiguous and film.dtype == np.float64 and film.shape == (h, w, 3)
        self._ck(self.lib.sppm_resolve(self.h, cnt, onetime, film.ctypes.data), "sppm_resolve")
        return film

    # ---- shared statistics across eye passes
    def shared_begin(self, w: int, h: int):
        self._ck(self.lib.sppm_shared_begin(self.h, w, h), "sppm_shared_begin")
        self._shared_wh = (w, h)

    def shared_load(self):
        self._ck(self.lib.sppm_shared_load(self.h), "sppm_shared_load")

    def shared_fold(self, photons_emitted: int):
        self._ck(self.lib.sppm_shared_fold(self.h, int(photons_emitted)), "sppm_shared_fold")

    def shared_resolve(self, out: np.ndarray | None = None) -> np.ndarray:
        w, h = self._shared_wh
        film = out if out is not None else np.empty((h, w, 3), dtype=np.float64)
        assert film.flags.c_contiguous and film.dtype == np.float64 and film.shape == (h, w, 3)
        self._ck(self.lib.sppm_shared_resolve(self.h, film.ctypes.data), "sppm_shared_resolve")
        return film

    def shared_export(self) -> dict:
        w, h = self._shared_wh
        r2, n, tau = np.empty((h, w)), np.empty((h, w)), np.empty((3, h, w))
        ph = C.c_uint64()
        self._ck(self.lib.sppm_shared_export(self.h, r2.ctypes.data, n.ctypes.data, tau.ctypes.data, C.byref(ph)),
                 "sppm_shared_export")
        return dict(r2=r2, n=n, tau=np.ascontiguousarray(tau.transpose(1, 2, 0)), photons=ph.value)

    # ---- JPEG decode / encode on the device (replaces cv::imread / cv::imwrite)
    def jpeg_decode(self, data: bytes) -> np.ndarray:
        buf = np.frombuffer(data, dtype=np.uint8)
        rows, cols = C.c_int32(), C.c_int32()
        self._ck(self.lib.sppm_jpeg_info(self.h, buf.ctypes.data, buf.size, C.byref(rows), C.byref(cols)), "sppm_jpeg_info")
        out = np.empty((rows.value, cols.value, 3), dtype=np.uint8)
        self._ck(self.lib.sppm_jpeg_decode(self.h, buf.ctypes.data, buf.size, rows.value, cols.value, out.ctypes.data), "sppm_jpeg_decode")
        return out

    def jpeg_encode(self, bgr8: np.ndarray, quality: int = 95) -> bytes:
        img = np.ascontiguousarray(bgr8, dtype=np.uint8)
        assert img.ndim == 3 and img.shape[2] == 3
        n = C.c_int64()
        self._ck(self.lib.sppm_jpeg_encode(self.h, img.ctypes.data, img.shape[0], img.shape[1], quality, None, 0, C.byref(n)), "sppm_jpeg_encode")
        out = np.empty(n.value + 4096, dtype=np.uint8)
        self._ck(self.lib.sppm_jpeg_encode(self.h, img.ctypes.data, img.shape[0], img.shape[1], quality, out.ctypes.data, out.size, C.byref(n)),
                 "sppm_jpeg_encode")
        return out[:n.value].tobytes()

    def tonemap(self, film: np.ndarray, div: float) -> np.ndarray:
        film = _f64(film)
        h, w = film.shape[:2]
        out = np.empty((h, w, 3), dtype=np.uint8)
        self._ck(self.lib.sppm_tonemap(self.h, film.ctypes.data, w, h, float(div), out.ctypes.data), "sppm_tonemap")
        return out

    def synchronize(self):
        self._ck(self.lib.sppm_device_synchronize(self.h), "sppm_device_synchronize")

    def timer_record(self, slot: int):
        self._ck(self.lib.sppm_timer_record(self.h, slot), "sppm_timer_record")

    def timer_elapsed_ms(self) -> float:
        ms = C.c_double()
        self._ck(self.lib.sppm_timer_elapsed_ms(self.h, C.byref(ms)), "sppm_timer_elapsed_ms")
        return ms.value

    # -- multi-GPU
    def comm_unique_id(self) -> bytes:
        buf = C.create_string_buffer(128)
        rc = self.lib.sppm_comm_unique_id(buf)
        if rc != 0:
            raise SppmError("sppm_comm_unique_id failed: libnccl.so.2 not loadable")
        return buf.raw

    def comm_init_rank(self, rank: int, nranks: int, uid: bytes):
        buf = C.create_string_buffer(uid, 128)
        self._ck(self.lib.sppm_comm_init_rank(self.h, rank, nranks, buf), "sppm_comm_init_rank")

    def accum_device_ptrs(self):
        a, b, n = C.c_void_p(), C.c_void_p(), C.c_int64()
        self._ck(self.lib.sppm_accum_device_ptrs(self.h, C.byref(a), C.byref(b), C.byref(n)), "sppm_accum_device_ptrs")
        return a.value, b.value, n.value

    # -- hooks
    def intersect_batch(self, rays6, ids=None, pass_id=0, node=0) -> np.ndarray:
        rays6 = _f64(rays6, (-1, 6))
        n = rays6.shape[0]
        out = np.zeros(n, dtype=hit_dtype)
        idp = None
        if ids is not None:
            ids = np.ascontiguousarray(ids, dtype=np.uint64)
            idp = ids.ctypes.data
        self._ck(self.lib.sppm_intersect_batch(self.h, rays6.ctypes.data, n, idp, pass_id, node, out.ctypes.data),
                 "sppm_intersect_batch")
        return out

    def bezier_patch_batch(self, rays6, patch_idx, seeds_uv=None) -> np.ndarray:
        rays6 = _f64(rays6, (-1, 6))
        n = rays6.shape[0]
        patch_idx = np.ascontiguousarray(patch_idx, dtype=np.int32)
        out = np.empty((n, 3), dtype=np.float64)
        sp = None
        if seeds_uv is not None:
            seeds_uv = _f64(seeds_uv, (n, self.params.newton_restarts, 2))
            sp = seeds_uv.ctypes.data
        self._ck(self.lib.sppm_bezier_patch_batch(self.h, rays6.ctypes.data, patch_idx.ctypes.data, n, sp, out.ctypes.data),
                 "sppm_bezier_patch_batch")
        return out

    def bezier_patch_bench(self, n: int, reps: int, seed: int = 1) -> float:
        ms = C.c_double()
        self._ck(self.lib.sppm_bezier_patch_bench(self.h, n, reps, seed, C.byref(ms)), "sppm_bezier_patch_bench")
        return ms.value

    def fp64_peak(self) -> float:
        tf = C.c_double()
        self._ck(self.lib.sppm_fp64_peak(self.h, C.byref(tf)), "sppm_fp64_peak")
        return tf.value

    def import_hitpoints(self, pos3, n3, wgt3, r2, pix, w, h):
        pos3, n3, wgt3 = _f64(pos3, (-1, 3)), _f64(n3, (-1, 3)), _f64(wgt3, (-1, 3))
        r2 = _f64(r2, (-1,))
        pix = np.ascontiguousarray(pix, dtype=np.int32).reshape(-1, 2)
        n = pos3.shape[0]
        self._ck(self.lib.sppm_import_hitpoints(self.h, n, _ptr(pos3), _ptr(n3), _ptr(wgt3), _ptr(r2), _ptr(pix), w, h),
                 "sppm_import_hitpoints")

    def export_hitpoints(self):
        n = C.c_int64()
        self._ck(self.lib.sppm_export_hitpoints(self.h, 0, None, None, None, None, None, None, C.byref(n)), "sppm_export_hitpoints")
        n = n.value
        pos3, n3, wgt3 = (np.empty((n, 3)) for _ in range(3))
        r2 = np.empty(n)
        pix = np.empty((n, 2), dtype=np.int32)
        node = np.empty(n, dtype=np.uint32)
        m = C.c_int64()
        self._ck(self.lib.sppm_export_hitpoints(self.h, n, _ptr(pos3), _ptr(n3), _ptr(wgt3), _ptr(r2), _ptr(pix), _ptr(node),
                                                C.byref(m)), "sppm_export_hitpoints")
        return dict(pos=pos3, n=n3, wgt=wgt3, r2=r2, pix=pix, node=node)

    def gather_deposits(self, x3, n3, flux3):
        x3, n3, flux3 = _f64(x3, (-1, 3)), _f64(n3, (-1, 3)), _f64(flux3, (-1, 3))
        self._ck(self.lib.sppm_gather_deposits(self.h, x3.shape[0], _ptr(x3), _ptr(n3), _ptr(flux3)), "sppm_gather_deposits")

    def export_stats(self, n: int):
        newN, N = np.zeros(n, dtype=np.int32), np.zeros(n, dtype=np.int32)
        r2, tao = np.zeros(n), np.zeros((n, 3))
        self._ck(self.lib.sppm_export_stats(self.h, n, _ptr(newN), _ptr(N), _ptr(r2), _ptr(tao)), "sppm_export_stats")
        return dict(newN=newN, N=N, r2=r2, tao=tao)

    def keep_deposits(self, keep_all: bool):
        self._ck(self.lib.sppm_keep_deposits(self.h, 1 if keep_all else 0), "sppm_keep_deposits")

    def export_deposits(self):
        m = C.c_int64()
        self._ck(self.lib.sppm_export_deposits(self.h, 0, None, None, None, None, None, C.byref(m)), "sppm_export_deposits")
        m = m.value
        x3, n3, f3 = (np.empty((m, 3)) for _ in range(3))
        pid = np.empty(m, dtype=np.uint64)
        bounce = np.empty(m, dtype=np.int32)
        mm = C.c_int64()
        self._ck(self.lib.sppm_export_deposits(self.h, m, _ptr(x3), _ptr(n3), _ptr(f3), _ptr(pid), _ptr(bounce), C.byref(mm)),
                 "sppm_export_deposits")
        return dict(x=x3, n=n3, flux=f3, photon=pid, bounce=bounce)

    def counters(self) -> dict:
        c = abi.SppmCounters()
        self._ck(self.lib.sppm_get_counters(self.h, C.byref(c)), "sppm_get_counters")
        return c.as_dict()

    def set_profiling(self, on: bool):
        """Serialised batches with an event between launches: per-kernel ms in counters() (measurement only)."""
        self._ck(self.lib.sppm_set_profiling(self.h, 1 if on else 0), "sppm_set_profiling")

    def reset_counters(self):
        self._ck(self.lib.sppm_reset_counters(self.h), "sppm_reset_counters")


hit_dtype = np.dtype([("t", "f8"), ("pos", "f8", 3), ("n", "f8", 3), ("color", "f8", 3), ("u", "f8"), ("v", "f8"),
                      ("obj", "i4"), ("patch", "i4"), ("inside", "i4"), ("pad", "i4")])
assert hit_dtype.itemsize == C.sizeof(abi.SppmHit)


class RayTracer:
    """Raytracer.h:82-419.  `rt.camera = cam; rt.scene = scene; rt.render()` as in the
    reference (Raytracer.h:439-445); the film lands in `camera.pic`.

    cnt / onetime default to the reference's literals (Raytracer.h:218, 384).  With
    world > 1 every rank traces its share of each call's photons and the accumulators are
    summed over ranks (NCCL) before HitPoint::update.

    Photon k of round i is photon i*onetime + k of pass `pass_base` in the counter RNG.  While
    HitPoint::update is the reference's (a no-op for the hit points of a fresh eye pass, N = 0,
    Raytracer.h:70, Q1) the rounds are independent and `fuse_rounds` traces them as one photon
    pass of cnt*onetime photons (batches of at most 8 Mi photons over the library's lanes) instead of cnt
    launch sequences of onetime = 100 photons: same photons, same newN, tao in another summation
    order.  PROGRESSIVE updates keep one pass per round."""

    def __init__(self, ctx: Context | None = None, device: int = 0, rank: int = 0, world: int = 1):
        self.ctx = ctx if ctx is not None else Context(device)
        self.camera: Camera | None = None
        self.scene: Scene | None = None
        self.cnt = 10000
        self.onetime = 100
        self.pass_base = 0
        self.fuse_rounds = True
        self.rank, self.world = rank, world
        self.n_hitpoints = 0
        self._scene_uploaded = None

    def GetHitPoint(self):  # Raytracer.h:102-116
        if self._scene_uploaded is not self.scene:
            self.ctx.upload_scene(self.scene)
            self._scene_uploaded = self.scene
        self.n_hitpoints = self.ctx.eye_pass(self.camera, self.pass_base)
        return self.n_hitpoints

    def PhotonMap(self, cnt: int):  # Raytracer.h:210-295
        fuse = self.fuse_rounds and self.ctx.params.update_mode == abi.SPPM_UPDATE_REFERENCE
        for begin, count in photon_calls(cnt, self.onetime, fuse, self.rank, self.world):
            self.ctx.photon_pass(begin, count, self.pass_base)
            if self.world > 1:
                self.ctx.allreduce()
            self.ctx.update()
        self.ctx.resolve(cnt, self.onetime, self.camera.w, self.camera.h, out=self.camera.pic)

    def render(self):  # Raytracer.h:366-387
        self.GetHitPoint()
        self.ctx.build_index()
        self.PhotonMap(self.cnt)


class SPPMRayTracer:
    """Raytracer.h:421-477: outer loop over jittered replicas, film accumulation, tone map.
    The reference loops 100000 times; `sppmtime` bounds it here.  One device runs the
    `replicas` RayTracers of an iteration back to back (the reference's 4 OpenMP threads)."""

    def __init__(self, ctx: Context | None = None, device: int = 0):
        self.ctx = ctx if ctx is not None else Context(device)
        self.sppmtime = 1
        self.replicas = 4
        self.cnt = 10000
        self.onetime = 100
        self.jitter = 0.00015
        self.image = None
        # False: the reference's estimator, the average of independent PPM renders
        # (Raytracer.h:429-458).  True: one set of per-pixel statistics shared by all eye passes
        # (SPPM proper); camera.pic then holds the estimate itself.
        self.shared_statistics = False
        self.fuse_rounds = True  # see RayTracer

    def render(self, scene: Scene, camera: Camera):
        from .rng import unit_rand_vec, P_JITTER
        seed = self.ctx.params.seed
        if self.shared_statistics:
            return self._render_shared(scene, camera)
        for i in range(self.sppmtime):
            for j in range(self.replicas):
                c = camera.copy()
                if self.jitter:
                    c.MoveTo(camera.pos + self.jitter * unit_rand_vec(seed, i, j, P_JITTER))
                c.lookAt(camera.look, 1.)
                rt = RayTracer(self.ctx)
                rt._scene_uploaded = scene if (i or j) else None
                rt.camera, rt.scene, rt.cnt, rt.onetime = c, scene, self.cnt, self.onetime
                rt.pass_base = i * self.replicas + j
                rt.fuse_rounds = self.fuse_rounds
                rt.render()
                camera.pic += c.pic  # Raytracer.h:449-458
            self.image = self.ctx.tonemap(camera.pic, (i + 1) * float(self.replicas))  # Raytracer.h:460-470
        return self.image

    def _render_shared(self, scene: Scene, camera: Camera):
        from .rng import unit_rand_vec, P_JITTER
        ctx, seed = self.ctx, self.ctx.params.seed
        if ctx.params.update_mode != abi.SPPM_UPDATE_REFERENCE:
            raise ValueError("shared statistics: the per-hit-point update must stay the reference's (update_mode REFERENCE)")
        ctx.upload_scene(scene)
        ctx.shared_begin(camera.w, camera.h)
        n_light = len(scene.plvec)
        for i in range(self.sppmtime):
            for j in range(self.replicas):
                c = camera.copy()
                if self.jitter:
                    c.MoveTo(camera.pos + self.jitter * unit_rand_vec(seed, i, j, P_JITTER))
                c.lookAt(camera.look, 1.)
                pass_base = i * self.replicas + j
                ctx.eye_pass(c, pass_base)
                ctx.build_index()
                ctx.shared_load()
                for begin, count in photon_calls(self.cnt, self.onetime, self.fuse_rounds):
                    ctx.photon_pass(begin, count, pass_base)
                    ctx.update()
                ctx.shared_fold(self.cnt * self.onetime * n_light)
            ctx.shared_resolve(out=camera.pic)
            self.image = ctx.tonemap(camera.pic, 1.0)
        return self.image
