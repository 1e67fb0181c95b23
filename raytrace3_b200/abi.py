"""ctypes mirror of include/sppm_b200.h (PODs + prototypes) and the library loader.

The product path has NO CPU fallback: if libsppm_b200.so is missing it is built with nvcc
(raytrace3_b200.build); if that fails, or no sm_100 device is present when a context is
created, an exception is raised.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SPPM_B200_LIB") or os.path.join(_HERE, "lib", "libsppm_b200.so")

c_double3 = C.c_double * 3

SPPM_OBJ_PLANE, SPPM_OBJ_SPHERE, SPPM_OBJ_BEZIER = 0, 1, 2
SPPM_UPDATE_REFERENCE, SPPM_UPDATE_PROGRESSIVE = 0, 1
SPPM_NEWTON_EXACT, SPPM_NEWTON_FAST, SPPM_NEWTON_FAST_BERNSTEIN = 0, 1, 2
SPPM_TRACE_WAVEFRONT, SPPM_TRACE_PERSISTENT = 0, 1
SPPM_GATHER_EXACT, SPPM_GATHER_FILTERED = 0, 1
SPPM_MATH_EXACT, SPPM_MATH_FAST = 0, 1
SPPM_DOMAIN_PHOTON, SPPM_DOMAIN_EYE, SPPM_DOMAIN_TEST = 2, 3, 5


class SppmMaterial(C.Structure):
    _fields_ = [("diff", c_double3), ("refl", c_double3), ("refr", c_double3), ("refrn", C.c_double)]


class SppmObject(C.Structure):
    _fields_ = [
        ("kind", C.c_int32), ("tex_id", C.c_int32), ("first_patch", C.c_int32), ("n_patch", C.c_int32),
        ("mat", SppmMaterial), ("color", c_double3), ("p0", c_double3), ("norm", c_double3),
        ("radius", C.c_double), ("texU", c_double3), ("texV", c_double3),
    ]


class SppmLight(C.Structure):
    _fields_ = [("pos", c_double3), ("color", c_double3)]


class SppmTexture(C.Structure):
    _fields_ = [("rows", C.c_int32), ("cols", C.c_int32), ("bgr8", C.c_void_p)]


class SppmSceneDesc(C.Structure):
    _fields_ = [
        ("n_objects", C.c_int32), ("n_patches", C.c_int32), ("n_lights", C.c_int32), ("n_textures", C.c_int32),
        ("objects", C.POINTER(SppmObject)), ("patches", C.c_void_p),
        ("lights", C.POINTER(SppmLight)), ("textures", C.POINTER(SppmTexture)),
    ]


class SppmCameraDesc(C.Structure):
    _fields_ = [("pos", c_double3), ("dir", c_double3), ("du", c_double3), ("dv", c_double3),
                ("w", C.c_int32), ("h", C.c_int32)]


class SppmParams(C.Structure):
    _fields_ = [
        ("init_r2", C.c_double), ("alpha", C.c_double), ("max_dep", C.c_int32), ("newton_iters", C.c_int32),
        ("newton_restarts", C.c_int32), ("update_mode", C.c_int32), ("newton_mode", C.c_int32),
        ("trace_mode", C.c_int32), ("seed", C.c_uint64), ("gather_mode", C.c_int32), ("count_pairs", C.c_int32), ("math_mode", C.c_int32), ("tex_lod", C.c_int32),
    ]


class SppmHit(C.Structure):
    _fields_ = [
        ("t", C.c_double), ("pos", c_double3), ("n", c_double3), ("color", c_double3),
        ("u", C.c_double), ("v", C.c_double),
        ("obj", C.c_int32), ("patch", C.c_int32), ("inside", C.c_int32), ("pad", C.c_int32),
    ]


class SppmCounters(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in (
        "photons", "isect_queries", "patch_solves", "deposits", "deposits_kept", "cand_pairs", "acc_pairs",
        "hit_points", "grid_cells", "launches")] + [(n, C.c_double) for n in (
            "ms_eye", "ms_build", "ms_trace", "ms_sort", "ms_gather", "ms_update", "ms_resolve")] + [
                ("patch_culled", C.c_uint64), ("spilled_photons", C.c_uint64)] + [(n, C.c_double) for n in (
                    "ms_step", "ms_newton", "ms_spill", "ms_bin_scan", "ms_bin_scatter", "ms_gather_prep", "ms_gather_kernel",
                    "ms_allreduce")]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


def default_params() -> SppmParams:
    """The reference's literals; identical to sppm_default_params()."""
    return SppmParams(init_r2=2.0, alpha=0.7, max_dep=13, newton_iters=10, newton_restarts=50,
                      update_mode=SPPM_UPDATE_REFERENCE, newton_mode=SPPM_NEWTON_EXACT, trace_mode=0,
                      seed=0x5EED5EED, gather_mode=SPPM_GATHER_FILTERED, count_pairs=1, math_mode=SPPM_MATH_EXACT, tex_lod=0)


VP, I32, I64, U64, DBL = C.c_void_p, C.c_int32, C.c_int64, C.c_uint64, C.c_double

# name -> (restype, argtypes); every symbol declared in include/sppm_b200.h
PROTOTYPES = {
    "sppm_abi_version": (C.c_int, []),
    "sppm_ctx_create": (C.c_int, [C.c_int, C.POINTER(VP)]),
    "sppm_ctx_destroy": (None, [VP]),
    "sppm_last_error": (C.c_char_p, [VP]),
    "sppm_default_params": (C.c_int, [C.POINTER(SppmParams)]),
    "sppm_set_params": (C.c_int, [VP, C.POINTER(SppmParams)]),
    "sppm_scene_upload": (C.c_int, [VP, C.POINTER(SppmSceneDesc)]),
    "sppm_eye_pass": (C.c_int, [VP, C.POINTER(SppmCameraDesc), U64, C.POINTER(I64)]),
    "sppm_eye_pass_rows": (C.c_int, [VP, C.POINTER(SppmCameraDesc), U64, I32, I32, I32, C.POINTER(I64)]),
    "sppm_allgather_hitpoints": (C.c_int, [VP, C.POINTER(I64)]),
    "sppm_build_index": (C.c_int, [VP]),
    "sppm_photon_pass": (C.c_int, [VP, U64, U64, U64]),
    "sppm_allreduce": (C.c_int, [VP]),
    "sppm_update": (C.c_int, [VP]),
    "sppm_resolve": (C.c_int, [VP, I32, I32, VP]),
    "sppm_tonemap": (C.c_int, [VP, VP, I32, I32, DBL, VP]),
    "sppm_shared_begin": (C.c_int, [VP, I32, I32]),
    "sppm_shared_load": (C.c_int, [VP]),
    "sppm_shared_fold": (C.c_int, [VP, U64]),
    "sppm_shared_resolve": (C.c_int, [VP, VP]),
    "sppm_shared_export": (C.c_int, [VP, VP, VP, VP, C.POINTER(U64)]),
    "sppm_shared_import": (C.c_int, [VP, VP, VP, VP, U64]),
    "sppm_jpeg_info": (C.c_int, [VP, VP, I64, C.POINTER(I32), C.POINTER(I32)]),
    "sppm_jpeg_decode": (C.c_int, [VP, VP, I64, I32, I32, VP]),
    "sppm_jpeg_encode": (C.c_int, [VP, VP, I32, I32, I32, VP, I64, C.POINTER(I64)]),
    "sppm_comm_unique_id": (C.c_int, [VP]),
    "sppm_comm_init_rank": (C.c_int, [VP, C.c_int, C.c_int, VP]),
    "sppm_accum_device_ptrs": (C.c_int, [VP, C.POINTER(VP), C.POINTER(VP), C.POINTER(I64)]),
    "sppm_device_synchronize": (C.c_int, [VP]),
    "sppm_stream_handle": (VP, [VP]),
    "sppm_timer_record": (C.c_int, [VP, C.c_int]),
    "sppm_timer_elapsed_ms": (C.c_int, [VP, C.POINTER(DBL)]),
    "sppm_intersect_batch": (C.c_int, [VP, VP, I64, VP, U64, I32, VP]),
    "sppm_bezier_patch_batch": (C.c_int, [VP, VP, VP, I64, VP, VP]),
    "sppm_bezier_patch_bench": (C.c_int, [VP, I64, I32, U64, C.POINTER(DBL)]),
    "sppm_fp64_peak": (C.c_int, [VP, C.POINTER(DBL)]),
    "sppm_import_hitpoints": (C.c_int, [VP, I64, VP, VP, VP, VP, VP, I32, I32]),
    "sppm_export_hitpoints": (C.c_int, [VP, I64, VP, VP, VP, VP, VP, VP, C.POINTER(I64)]),
    "sppm_gather_deposits": (C.c_int, [VP, I64, VP, VP, VP]),
    "sppm_export_stats": (C.c_int, [VP, I64, VP, VP, VP, VP]),
    "sppm_keep_deposits": (C.c_int, [VP, I32]),
    "sppm_export_deposits": (C.c_int, [VP, I64, VP, VP, VP, VP, VP, C.POINTER(I64)]),
    "sppm_get_counters": (C.c_int, [VP, C.POINTER(SppmCounters)]),
    "sppm_reset_counters": (C.c_int, [VP]),
    "sppm_set_profiling": (C.c_int, [VP, I32]),
    "sppm_host_pin": (C.c_int, [VP, U64]),
    "sppm_host_unpin": (C.c_int, [VP]),
}

_lib = None


def load_library(build_if_missing: bool = True) -> C.CDLL:
    """dlopen libsppm_b200.so (building it in-tree first if absent). Raises on failure."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        if not build_if_missing:
            raise RuntimeError(f"{LIB_PATH} is missing (run `python -m raytrace3_b200.build`); there is no CPU fallback")
        from . import build as _build
        _build.build()
    lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib
