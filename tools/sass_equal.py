#!/usr/bin/env python
"""Is the device code of the working tree the same as that of a git revision?

    python tools/sass_equal.py <rev>        # e.g. the last commit whose GPU tests ran on a B200

Builds libsppm_b200.so of <rev> in a temporary git worktree (nvcc cross-compiles sm_100a without a
GPU), dumps the SASS of every object of both builds with cuobjdump and compares it after removing
what legitimately differs between two checkouts: the -lineinfo file markers, the source path
identifier, the hash nvcc puts into the names of anonymous-namespace kernels and the order in which
the kernels are defined.  Exit status 0 =
identical instruction streams, i.e. a refactor (moving code between files, host-only changes) is
proven device-neutral and the GPU evidence of <rev> still describes the tree; 1 = they differ.
"""
import hashlib
import os
import re
import shutil
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OBJECTS = ["abi", "trace", "wavefront", "gather", "imageio"]


def sass(obj_path: str) -> str:
    out = subprocess.run(["cuobjdump", "-sass", obj_path], capture_output=True, text=True, check=True).stdout
    keep = [l for l in out.splitlines() if "//##" not in l and "identifier =" not in l]
    text = re.sub(r"_GLOBAL__N__[0-9a-f]+_", "_G_", "\n".join(keep))
    # one block per kernel, sorted by name: the order of definition in the source is not device code
    parts = re.split(r"(?m)^(?=\s*Function : )", text)
    return parts[0] + "".join(sorted(parts[1:]))


def build(tree: str) -> str:
    subprocess.run([sys.executable, "-c", "from raytrace3_b200 import build; build.build()"], cwd=tree, check=True,
                   capture_output=True, env=dict(os.environ, PYTHONPATH=tree))
    return os.path.join(tree, "raytrace3_b200", "lib")


def main() -> int:
    if len(sys.argv) != 2:
        print(__doc__)
        return 2
    rev = sys.argv[1]
    tmp = tempfile.mkdtemp(prefix="sass_equal_")
    wt = os.path.join(tmp, "wt")
    try:
        subprocess.run(["git", "worktree", "add", "-q", "--detach", wt, rev], cwd=ROOT, check=True)
        old, new = build(wt), build(ROOT)
        same = True
        for o in OBJECTS:
            a, b = (sass(os.path.join(d, o + ".o")) for d in (old, new))
            eq = a == b
            same &= eq
            print("%-10s %s  %s %s" % (o, "identical" if eq else "DIFFERENT", hashlib.md5(a.encode()).hexdigest()[:12],
                                       hashlib.md5(b.encode()).hexdigest()[:12]))
        return 0 if same else 1
    finally:
        subprocess.run(["git", "worktree", "remove", "--force", wt], cwd=ROOT, capture_output=True)
        subprocess.run(["git", "worktree", "prune"], cwd=ROOT, capture_output=True)
        shutil.rmtree(tmp, ignore_errors=True)


if __name__ == "__main__":
    sys.exit(main())
