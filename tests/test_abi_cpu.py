"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol that
include/sppm_b200.h declares, the ctypes mirror matches the header's struct sizes, and the
product path fails loudly (no CPU fallback) when no B200 is present."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

from raytrace3_b200 import abi
from raytrace3_b200.api import photon_range

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "sppm_b200.h")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(sppm_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = abi.load_library()
    names = declared_symbols()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), f"{n} declared in sppm_b200.h but not exported"
        assert n in abi.PROTOTYPES, f"{n} has no ctypes prototype"
    assert set(abi.PROTOTYPES) == set(names)
    assert lib.sppm_abi_version() == 3


def test_struct_layout_matches_header(tmp_path):
    """sizeof of every POD as the C compiler sees it == the ctypes mirror."""
    names = ["SppmMaterial", "SppmObject", "SppmLight", "SppmTexture", "SppmSceneDesc", "SppmCameraDesc", "SppmParams",
             "SppmHit", "SppmCounters"]
    prog = '#include <stdio.h>\n#include "sppm_b200.h"\nint main(){' + "".join(
        f'printf("%zu\\n", sizeof({n}));' for n in names) + "return 0;}"
    src = tmp_path / "sz.c"
    src.write_text(prog)
    exe = tmp_path / "sz"
    subprocess.run(["/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc", "-I", os.path.dirname(HEADER), str(src), "-o", str(exe)], check=True)
    sizes = [int(x) for x in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    for n, sz in zip(names, sizes):
        assert C.sizeof(getattr(abi, n)) == sz, n


def test_default_params_are_the_reference_literals():
    lib = abi.load_library()
    p = abi.SppmParams()
    assert lib.sppm_default_params(C.byref(p)) == 0
    q = abi.default_params()
    for f, _ in abi.SppmParams._fields_:
        assert getattr(p, f) == getattr(q, f), f
    assert (p.init_r2, p.alpha, p.max_dep, p.newton_iters, p.newton_restarts) == (2.0, 0.7, 13, 10, 50)


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    lib = abi.load_library()
    h = C.c_void_p()
    rc = lib.sppm_ctx_create(0, C.byref(h))
    assert rc == -3 and not h.value  # SPPM_ERR_NO_DEVICE
    assert b"no CUDA device" in lib.sppm_last_error(None)
    from raytrace3_b200 import Context, SppmError
    with pytest.raises(SppmError):
        Context(0)


def test_null_and_bad_arguments_are_rejected():
    lib = abi.load_library()
    assert lib.sppm_default_params(None) == -1
    assert lib.sppm_ctx_create(0, None) == -1
    assert lib.sppm_set_params(None, None) == -1
    assert lib.sppm_photon_pass(None, 0, 1, 0) == -1
    assert lib.sppm_get_counters(None, None) == -1
    lib.sppm_ctx_destroy(None)  # no-op


def test_photon_range_partitions_exactly():
    for onetime in (1, 7, 100, 2_000_000, 16_000_001):
        for world in (1, 2, 3, 4, 8):
            spans = [photon_range(onetime, r, world) for r in range(world)]
            assert spans[0][0] == 0
            for (b0, c0), (b1, _) in zip(spans, spans[1:]):
                assert b0 + c0 == b1
            assert spans[-1][0] + spans[-1][1] == onetime
            assert max(c for _, c in spans) - min(c for _, c in spans) <= 1


def test_photon_calls_cover_the_eye_pass_exactly(tmp_path):
    """RayTracer::PhotonMap's round fusion: whatever the fusion and the number of ranks, the photon
    passes issued for PhotonMap(cnt) cover the ids [0, cnt*onetime) exactly once; the reference's
    defaults (10000 rounds x 100 photons) become one pass; the C++ twin agrees."""
    from raytrace3_b200 import photon_calls, rounds_per_call
    from raytrace3_b200.api import FUSE_PHOTONS
    assert photon_calls(10000, 100, True) == [(0, 1_000_000)]
    assert photon_calls(3, 100, False) == [(0, 100), (100, 100), (200, 100)]
    assert len(photon_calls(1000, 64_000_000, True)) == 250      # whole rounds, <= 2^28 photons per call
    for cnt, onetime in ((1, 1), (7, 13), (10000, 100), (10, 2_000_000), (5, 300_000_001), (33, 64_000_000)):
        for fuse in (False, True):
            for world in (1, 2, 3, 8):
                spans = sorted(sp for r in range(world) for sp in photon_calls(cnt, onetime, fuse, r, world) if sp[1] > 0)
                pos = 0
                for b, c in spans:
                    assert b == pos
                    pos += c
                assert pos == cnt * onetime
                per_rank = photon_calls(cnt, onetime, fuse, 0, world)
                assert len(per_rank) == -(-cnt // rounds_per_call(onetime, fuse))
                assert all(c <= max(FUSE_PHOTONS, onetime) for _, c in per_rank)
                if not fuse:
                    assert len(per_rank) == cnt
    vals = [1, 7, 100, 2_000_000, 64_000_000, 300_000_001, 1 << 28, (1 << 28) + 1]
    prog = ('#include <cstdio>\n#include "SppmBridge.h"\nint main(){' +
            "".join(f'std::printf("%llu %llu\\n", (unsigned long long)sppm_host::rounds_per_call({v}ull, true), '
                    f'(unsigned long long)sppm_host::rounds_per_call({v}ull, false));' for v in vals) + "return 0;}")
    src = tmp_path / "rpc.cpp"
    src.write_text(prog)
    exe = tmp_path / "rpc"
    inc = os.path.join(os.path.dirname(os.path.dirname(HEADER)), "raytrace3_b200", "include", "raytracer")
    subprocess.run(["/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++", "-std=c++17", "-I", inc, str(src), "-o", str(exe)], check=True)
    out = [tuple(int(x) for x in ln.split()) for ln in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.splitlines()]
    assert out == [(rounds_per_call(v, True), rounds_per_call(v, False)) for v in vals]


def test_scene_flatten_roundtrip():
    from raytrace3_b200.scene import make_scene
    sc = make_scene("reference", textures="synthetic")
    d = sc.flatten()
    assert (d.n_objects, d.n_patches, d.n_lights, d.n_textures) == (9, 32, 1, 3)
    assert [d.objects[i].kind for i in range(9)] == [0, 0, 0, 0, 0, 1, 1, 1, 2]
    assert [d.objects[i].tex_id for i in range(9)] == [0, 0, -1, 1, 0, -1, -1, 2, -1]  # Scene.h:134-141 (ids in first-use order)
    assert d.objects[8].first_patch == 0 and d.objects[8].n_patch == 32
    assert list(d.objects[3].norm) == [0.0, -1.0, 0.0]  # Q7: bottom plane normal points down
    assert d.objects[6].mat.refrn == 1.5
    P = np.ctypeslib.as_array(C.cast(d.patches, C.POINTER(C.c_double)), shape=(32, 16, 3))
    assert np.array_equal(P, sc.objvec[8].P)
    assert d.textures[0].rows == 2094 and d.textures[0].cols == 2950 and d.textures[2].rows == 399
