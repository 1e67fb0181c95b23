"""Scene description files: the shipped reference.scene reproduces the hard-coded
scene of Scene::init (Scene.h:116-157) and camera of main.cpp:24-27 exactly."""
import ctypes as C
import os
import shutil

import numpy as np
import pytest

from raytrace3_b200 import abi
from raytrace3_b200.scene import make_scene, reference_camera, teapot_patches, write_bpt
from raytrace3_b200.scenefile import SceneFileError, load_scene_file

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SCENE = os.path.join(ROOT, "examples", "scenes", "reference.scene")


def _stage(tmp_path):
    shutil.copy(SCENE, tmp_path / "reference.scene")
    write_bpt(tmp_path / "teapot.bpt", teapot_patches())
    return str(tmp_path / "reference.scene")


def test_reference_scene_file_equals_hard_coded_scene(tmp_path):
    sc, job = load_scene_file(_stage(tmp_path))
    ref = make_scene("reference", textures=None)
    a, b = sc.flatten(), ref.flatten()
    assert (a.n_objects, a.n_patches, a.n_lights, a.n_textures) == (b.n_objects, b.n_patches, b.n_lights, 0)
    n = a.n_objects * C.sizeof(abi.SppmObject)
    assert C.string_at(a.objects, n) == C.string_at(b.objects, n)
    np.testing.assert_array_equal(sc._keep[1], ref._keep[1])          # control points after the ctor transform
    assert C.string_at(a.lights, C.sizeof(abi.SppmLight)) == C.string_at(b.lights, C.sizeof(abi.SppmLight))
    cam, rc = job.camera(), reference_camera(320, 240)
    for k in ("pos", "dir", "du", "dv"):
        np.testing.assert_array_equal(getattr(cam, k), getattr(rc, k))
    assert (job.w, job.h, job.passes, job.replicas, job.rounds, job.photons) == (320, 240, 1, 4, 10, 100000)
    assert job.seed == 0x5EED5EED and not job.shared and job.format == "ppm"
    p = job.params()
    assert p.newton_mode == abi.SPPM_NEWTON_EXACT and p.gather_mode == abi.SPPM_GATHER_FILTERED


def test_scene_file_options_and_errors(tmp_path):
    f = tmp_path / "s.scene"
    f.write_text("""film 64 48
camera 0 0 10  0 0 0  1
material M diff .5 .5 .5
material G refr .9 .9 .9 refrn 1.33
plane 0 0 0  0 0 1  1 1 1  M
sphere 2  0 0 3  .9 .9 .9  G
light 0 5 5  100 100 100
render passes 3 rounds 2 photons 500 seed 0x10 shared 1 newton fast gather exact math fast jitter 0
output prefix out format jpg quality 80 checkpoint ck.bin every 2
""")
    sc, job = load_scene_file(str(f))
    assert len(sc.objvec) == 2 and sc.objvec[1].material.refrn == 1.33 and sc.objvec[1].radius == 2
    assert job.shared and job.passes == 3 and job.seed == 16 and job.jitter == 0 and job.quality == 80
    assert (job.newton_mode, job.gather_mode, job.math_mode) == (abi.SPPM_NEWTON_FAST, abi.SPPM_GATHER_EXACT, abi.SPPM_MATH_FAST)
    assert job.checkpoint == "ck.bin" and job.checkpoint_every == 2 and job.format == "jpg"
    for bad in ("plane 0 0 0 0 0 1 1 1 1 NOPE\nlight 0 0 0 1 1 1\n", "blob 1 2 3\n", "film 10\n",
                "material M diff 1 1 1\nplane 0 0 0 0 0 1 1 1 1 M\n"):
        f.write_text(bad)
        with pytest.raises(SceneFileError):
            load_scene_file(str(f))


def test_cpp_scene_loader_equals_python_twin(tmp_path):
    """The C++ host mirror (SceneFile.h -> Scene::flatten, Camera::desc) and the Python twin produce
    the same flattened scene from the same file, byte for byte: objects, control points after the
    Bezier3Obj constructor transform, lights, camera basis, job settings.  No GPU involved."""
    import subprocess
    from raytrace3_b200 import abi
    abi.load_library()
    r = subprocess.run(["make", "-C", os.path.join(ROOT, "examples"), "scene_dump"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    path = _stage(tmp_path)
    text = open(path).read().replace("render  passes 1", "render  passes 3 math fast").replace("film 320 240", "film 200 150")
    open(path, "w").write(text + "sphere 3.5  10 20 30  .1 .2 .3  REFR0\nlight 1 2 3  4 5 6\n")
    out = tmp_path / "dump.bin"
    r = subprocess.run([os.path.join(ROOT, "examples", "scene_dump"), path, str(out)], capture_output=True, text=True, cwd=tmp_path)
    assert r.returncode == 0, r.stdout + r.stderr
    raw = out.read_bytes()
    hdr = np.frombuffer(raw[:32], dtype=np.int64)
    sc, job = load_scene_file(path)
    d = sc.flatten()
    assert tuple(hdr) == (d.n_objects, d.n_patches, d.n_lights, 0)
    off = 32
    n = d.n_objects * C.sizeof(abi.SppmObject)
    assert raw[off:off + n] == C.string_at(d.objects, n)
    off += n
    P = sc._keep[1]
    assert raw[off:off + P.nbytes] == P.tobytes()
    off += P.nbytes
    n = d.n_lights * C.sizeof(abi.SppmLight)
    assert raw[off:off + n] == C.string_at(d.lights, n)
    off += n
    cd = job.camera().desc()
    n = C.sizeof(abi.SppmCameraDesc)
    assert raw[off:off + n] == C.string_at(C.addressof(cd), n)
    off += n
    jobv = np.frombuffer(raw[off:off + 80], dtype=np.int64)
    assert tuple(jobv) == (200, 150, 3, 4, 10, 100000, 0x5EED5EED, 0, job.newton_mode * 100 + job.gather_mode * 10 + job.math_mode, 95)
    assert job.math_mode == abi.SPPM_MATH_FAST and d.n_objects == 10 and d.n_lights == 2
