"""In-tree build of libsppm_b200.so (nvcc, sm_100a only) — `python -m raytrace3_b200.build`.

The library is compiled with -fmad=false: the reference (MSVC /fp:precise) never contracts
a*b+c, and bit-exact `newN` / EXACT-mode Newton depend on that; FMA appears only where the
kernels call fma() explicitly (FAST Newton).
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(ROOT, "csrc")
LIBDIR = os.environ.get("SPPM_BUILD_DIR") or os.path.join(ROOT, "lib")  # SPPM_BUILD_DIR + SPPM_NVCC_EXTRA: an A/B variant beside the product build
LIB = os.path.join(LIBDIR, "libsppm_b200.so")
SOURCES = ["abi.cu", "trace.cu", "wavefront.cu", "gather.cu", "imageio.cu"]
HEADERS = ["ctx.h", "scene.cuh", "scene_prep.h", "trace_kernels.cuh", "wavefront_kernels.cuh", "gather_kernels.cuh", "image_kernels.cuh", "device_math.cuh", os.path.join("..", "..", "include", "sppm_b200.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo", "-fmad=false",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-O2",
    "--expt-relaxed-constexpr",
    "-Xptxas", "-v", *(os.environ.get("SPPM_NVCC_EXTRA", "").split()),
]


def find_nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libsppm_b200.so cannot be built (there is no CPU fallback)")


def _stale(target: str, deps: list[str]) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(LIBDIR, exist_ok=True)
    nvcc = find_nvcc()
    hdrs = [os.path.join(CSRC, h) for h in HEADERS] + [os.path.abspath(__file__)]
    objs = []
    jobs = []
    for s in SOURCES:
        src = os.path.join(CSRC, s)
        obj = os.path.join(LIBDIR, s.replace(".cu", ".o"))
        objs.append(obj)
        if force or _stale(obj, [src] + hdrs):
            jobs.append([nvcc, *NVCC_FLAGS, "-I/usr/include", "-c", src, "-o", obj])

    def run(cmd):
        r = subprocess.run(cmd, capture_output=True, text=True)
        if verbose or r.returncode != 0:
            sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed: " + " ".join(cmd))
        return r.stderr

    logs = []
    if jobs:
        with ThreadPoolExecutor(max_workers=len(jobs)) as ex:
            logs = list(ex.map(run, jobs))
        with open(os.path.join(LIBDIR, "ptxas.log"), "w") as f:
            f.write("\n".join(logs))
    if jobs or force or _stale(LIB, objs):
        run([nvcc, "-shared", "-o", LIB, *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-lcudart", "-ldl"])
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
