"""API drop-in test: the body of the reference's main.cpp compiles against
the new same-named headers, and (on a GPU) renders the same film as the Python host mirror."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build():
    from raytrace3_b200 import abi
    abi.load_library()  # make sure libsppm_b200.so exists
    r = subprocess.run(["make", "-C", os.path.join(ROOT, "examples")], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    return os.path.join(ROOT, "examples", "main")


def test_reference_main_compiles_against_dropin_headers():
    exe = _build()
    assert os.path.exists(exe)
    # the example includes the reference's header names verbatim
    src = open(os.path.join(ROOT, "examples", "main.cpp")).read()
    for h in ("Vec3.h", "Obj.h", "Camera.h", "Element.h", "Light.h", "Scene.h", "Raytracer.h"):
        assert f'#include "{h}"' in src


@pytest.mark.gpu
def test_cpp_render_matches_python_mirror(tmp_path, ctx):
    from raytrace3_b200 import SPPMRayTracer, abi
    from raytrace3_b200.scene import make_scene, reference_camera, teapot_patches, write_bgr8, write_bpt
    exe = _build()
    sc = make_scene("reference", textures="synthetic")
    write_bpt(tmp_path / "teapot.bpt", teapot_patches())
    write_bgr8(tmp_path / "wall.bgr8", sc.objvec[0].texture)
    write_bgr8(tmp_path / "timg.bgr8", sc.objvec[3].texture)
    write_bgr8(tmp_path / "planet.bgr8", sc.objvec[7].texture)
    env = dict(os.environ, SPPM_ASSET_DIR=str(tmp_path), SPPM_W="48", SPPM_H="36", SPPM_ITERS="1", SPPM_CNT="2",
               SPPM_ONETIME="20000", SPPM_JITTER="0")
    r = subprocess.run([exe], capture_output=True, text=True, env=env, cwd=tmp_path, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    film_sum = float([l for l in r.stdout.splitlines() if l.startswith("film sum")][-1].split()[-1])
    ctx.set_params(abi.default_params())
    sp = SPPMRayTracer(ctx)
    sp.sppmtime, sp.cnt, sp.onetime, sp.jitter = 1, 2, 20000, 0.0
    cam = reference_camera(48, 36)
    sp.render(sc, cam)
    assert film_sum > 0
    np.testing.assert_allclose(film_sum, cam.pic.sum(), rtol=1e-9)
    assert (tmp_path / "aaanew0.ppm").exists()
    # shared per-pixel statistics across the 4 eye passes: C++ mirror == Python mirror
    r = subprocess.run([exe], capture_output=True, text=True, env=dict(env, SPPM_SHARED="1"), cwd=tmp_path, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    film_sum = float([l for l in r.stdout.splitlines() if l.startswith("film sum")][-1].split()[-1])
    sp.shared_statistics = True
    cam = reference_camera(48, 36)
    sp.render(sc, cam)
    np.testing.assert_allclose(film_sum, cam.pic.sum(), rtol=1e-9)


@pytest.mark.gpu
def test_cli_scene_file_checkpoint_resume_and_jpeg(tmp_path, ctx):
    """examples/sppm_render renders a scene description file (JPEG texture decoded
    by nvJPEG), resumes from a film checkpoint to the film of an uninterrupted run (both estimators), writes JPEG, and agrees with the Python twin on the same file."""
    cv2 = pytest.importorskip("cv2")
    import shutil
    from raytrace3_b200.scene import Texture, synthetic_texture, teapot_patches, write_bpt
    from raytrace3_b200.scenefile import load_scene_file, render_job
    _build()
    exe = os.path.join(ROOT, "examples", "sppm_render")
    shutil.copy(os.path.join(ROOT, "examples", "scenes", "reference.scene"), tmp_path / "ref.scene")
    write_bpt(tmp_path / "teapot.bpt", teapot_patches())
    ok, enc = cv2.imencode(".jpg", synthetic_texture(120, 160, 4), [cv2.IMWRITE_JPEG_QUALITY, 92])
    assert ok
    (tmp_path / "wall.jpg").write_bytes(enc.tobytes())          # timg.jpg / planet.jpg stay absent: untextured

    def run(*opts):
        r = subprocess.run([exe, "ref.scene", "--size", "48x36", "--replicas", "2", "--rounds", "1", "--photons", "20000", *opts],
                           capture_output=True, text=True, cwd=tmp_path, timeout=600)
        assert r.returncode == 0, r.stdout + r.stderr
        return [l for l in r.stdout.splitlines() if l.startswith("film sum")][-1].split()[-1]

    for mode in ([], ["--shared"]):
        straight = run("--passes", "2", *mode)
        run("--passes", "1", "--checkpoint", "ck.bin", *mode)
        resumed = run("--passes", "2", "--checkpoint", "ck.bin", "--resume", *mode)
        # same film after resume; the order in which a round's deposits are summed is not fixed
        # (atomic arrival order), so two runs agree to rounding, not bit for bit
        assert float(straight) > 0 and abs(float(resumed) - float(straight)) <= 1e-12 * float(straight)
    run("--passes", "1", "--format", "jpg", "--out", "shot")
    img = cv2.imread(str(tmp_path / "shot0.jpg"))
    assert img is not None and img.shape == (36, 48, 3) and img.mean() > 1
    # the Python twin on the same file (jitter 0: the two hosts draw their jitter differently)
    cpp = float(run("--passes", "1", "--jitter", "0"))
    sc, job = load_scene_file(str(tmp_path / "ref.scene"),
                              texture_loader=lambda p: Texture(ctx.jpeg_decode(open(p, "rb").read())))
    job.w, job.h, job.replicas, job.rounds, job.photons, job.passes, job.jitter = 48, 36, 2, 1, 20000, 1, 0.0
    film, _ = render_job(ctx, sc, job)
    np.testing.assert_allclose(cpp, film.sum(), rtol=1e-9)


def test_camera_film_is_one_block_with_the_reference_copy_semantics(tmp_path):
    """Camera.h:13 `Color** pic` of the drop-in: pic[i][j] addresses one double[h][w][3] block (Camera::film(), what
    sppm_resolve fills directly); a copy takes the pose and a fresh zeroed film (Camera.h:55-74); assigning a camera of
    another size re-allocates.  Runs without a GPU: page-locking is best effort."""
    from raytrace3_b200 import abi
    abi.load_library()
    src = tmp_path / "cam.cpp"
    src.write_text(r'''
#include <cstdio>
#include "Camera.h"
int main() {
  Camera a;
  a.init(7, 5);
  a.MoveTo(Vec3(50, 35, 230));
  a.lookAt(Vec3(50, 35.04, 229), 1.);
  if (!a.pic || a.film() != a.pic[0][0].val) return 1;
  for (int i = 0; i < a.h; i++)
    for (int j = 0; j < a.w; j++)
      for (int k = 0; k < 3; k++) {
        if (&a.pic[i][j].val[k] != a.film() + ((size_t)i * a.w + j) * 3 + k) return 2;
        if (a.pic[i][j].val[k] != 0) return 3;
        a.pic[i][j].val[k] = i * 100 + j * 10 + k;
      }
  Camera b = a;                       // pose copied, film fresh
  if (b.w != 7 || b.h != 5 || b.pic == a.pic || b.film() == a.film() || b.pic[4][6].val[2] != 0) return 4;
  if (b.pos.val[2] != 230 || b.du.val[0] != a.du.val[0]) return 5;
  Camera c;
  c.init(3, 2);
  c.lookAt(Vec3(0, 1, 0), 1.);
  c = a;                              // other size: re-allocated
  if (c.w != 7 || c.h != 5 || c.film() == nullptr) return 6;
  c.pic[4][6].val[2] = 1;             // the last element of the block is addressable
  a.clear();
  if (a.pic[4][6].val[2] != 0) return 7;
  std::puts("camera ok");
  return 0;
}
''')
    inc = os.path.join(ROOT, "raytrace3_b200", "include", "raytracer")
    lib = os.path.dirname(abi.LIB_PATH)
    exe = tmp_path / "cam"
    r = subprocess.run(["g++", "-O1", "-std=c++17", "-fopenmp", "-I", inc, str(src), "-o", str(exe), "-L", lib, "-lsppm_b200",
                        f"-Wl,-rpath,{lib}"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "camera ok" in r.stdout, (r.returncode, r.stdout, r.stderr)
