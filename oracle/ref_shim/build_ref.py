"""Build oracle/_ref/libref.so: the REFERENCE'S OWN headers (/root/reference/raytracer/*.h)
compiled against the thin OpenCV/FLANN shim in oracle/ref_shim/include, plus a C harness.
Only possible where /root/reference exists (the dev container); the built .so travels to the
GPU box with the snapshot.  The headers are read where they lie; three MSVC-only constructs are
patched in a scratch copy in a temporary directory outside the repo (removed after the compile;
no reference source text ever lands in the tree):
    Element.h:5   `extern class Obj;`            -> `class Obj;`
    Obj.h:47      `virtual ~Obj() = 0 {};`       -> `virtual ~Obj() {};`
    Bezier.h:236  `freopen("CON", "r", stdin);`  -> `freopen("/dev/null", "r", stdin);`
Everything else — every arithmetic statement — is the reference's code, unmodified."""
import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/raytracer"
OUT = os.path.join(HERE, "..", "_ref")
PATCHES = [
    ("Element.h", "extern class Obj;", "class Obj;"),
    ("Obj.h", "virtual ~Obj() = 0 {};", "virtual ~Obj() {};"),
    ("Bezier.h", 'freopen("CON", "r", stdin);', 'freopen("/dev/null", "r", stdin);'),
]


def main():
    if not os.path.isdir(REF):
        print("build_ref: /root/reference not present; keeping any prebuilt oracle/_ref/libref.so")
        return 0
    os.makedirs(OUT, exist_ok=True)
    shutil.rmtree(os.path.join(OUT, "src"), ignore_errors=True)    # older builds kept the scratch copy here
    src = tempfile.mkdtemp(prefix="sppm_ref_src_")
    try:
        return compile_ref(src)
    finally:
        shutil.rmtree(src, ignore_errors=True)


def compile_ref(src):
    for name in ("Vec3.h", "Element.h", "Obj.h", "Bezier.h", "Light.h", "Camera.h", "Scene.h", "Raytracer.h"):
        text = open(os.path.join(REF, name), encoding="latin-1").read()
        for f, a, b in PATCHES:
            if f == name:
                assert a in text, (f, a)
                text = text.replace(a, b)
        open(os.path.join(src, name), "w", encoding="latin-1").write(text)
    # the reference includes its third-party headers with backslashes (`#include <opencv2\\highgui.hpp>`, MSVC): g++ looks
    # the name up literally, so three forwarding headers with a backslash IN their file names satisfy them.  They are
    # written here, into the scratch directory, and not tracked (such names cannot be checked out on Windows).
    fwd = {"opencv2\\highgui.hpp": '#pragma once\n#include "cv_shim.hpp"\n',
           "flann\\flann.hpp": '#pragma once\n#include "flann_shim.hpp"\n',
           "Eigen\\Dense": "// Eigen is included by Bezier.h:3 but never used\n"}
    for name, text in fwd.items():
        open(os.path.join(src, name), "w").write(text)
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    cmd = [cxx, "-O2", "-std=c++17", "-ffp-contract=off", "-fno-fast-math", "-fPIC", "-shared", "-fopenmp", "-w",
           "-I", os.path.join(HERE, "include"), "-I", src, "-I", os.path.join(HERE, "..", "..", "include"),
           os.path.join(HERE, "ref_harness.cpp"), "-o", os.path.join(OUT, "libref.so")]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        return 1
    print("built", os.path.join(OUT, "libref.so"))
    return 0


if __name__ == "__main__":
    sys.exit(main())
