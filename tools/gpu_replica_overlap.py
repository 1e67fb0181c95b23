"""The reference's own loop shape (SPPMRayTracer: passes x 4 jittered replicas x (eye pass + 10000 rounds of 100
photons)) through examples/sppm_render, with the replicas of a pass on one context (--gpus 0) or spread over several
contexts of the SAME GPU (--gpus 0,0 / 0,0,0,0: one host thread each, so one replica's eye pass and grid build overlap
another's photon pass).  Prints the time per pass = (t(P2) - t(P1)) / (P2 - P1), which cancels start-up.
    gpurun --timeout 600 -- 'python tools/gpu_replica_overlap.py r2_m'
"""
import os, shutil, subprocess, sys, tempfile, time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from raytrace3_b200.scene import teapot_patches, write_bpt  # noqa: E402

tag = sys.argv[1] if len(sys.argv) > 1 else "r2_m"
out = os.path.join(ROOT, "gpurun_out")
os.makedirs(out, exist_ok=True)
w = tempfile.mkdtemp()
shutil.copy(os.path.join(ROOT, "examples", "scenes", "reference.scene"), os.path.join(w, "ref.scene"))
write_bpt(os.path.join(w, "teapot.bpt"), teapot_patches())
exe = os.path.join(ROOT, "examples", "sppm_render")
P1, P2 = 2, 12


def run(gpus, passes):
    t0 = time.perf_counter()
    r = subprocess.run([exe, "ref.scene", "--size", "1024x768", "--passes", str(passes), "--replicas", "4", "--rounds", "10000",
                        "--photons", "100", "--newton", "fast", "--math", "fast", "--gpus", gpus, "--out", os.path.join(w, "o")],
                       capture_output=True, text=True, cwd=w, timeout=300)
    dt = time.perf_counter() - t0
    if r.returncode != 0:
        raise SystemExit(f"--gpus {gpus} failed: {r.stdout[-400:]} {r.stderr[-400:]}")
    film = [l for l in r.stdout.splitlines() if l.startswith("film sum")]
    return dt, (film[-1] if film else "")


lines = []
for gpus in ("0", "0,0", "0,0,0,0", "0"):
    a, _ = run(gpus, P1)
    b, film = run(gpus, P2)
    per = (b - a) / (P2 - P1)
    lines.append(f"--gpus {gpus:8s} {P1} passes {a:6.2f} s, {P2} passes {b:6.2f} s -> {per * 1e3:7.1f} ms per pass of 4 replicas x "
                 f"(eye pass + 1 M photons) = {4e6 / per / 1e6:6.1f} M photons/s   [{film}]")
    print(lines[-1], flush=True)
open(os.path.join(out, f"{tag}_replica_overlap.txt"), "w").write("\n".join(lines) + "\n")
