#!/usr/bin/env bash
# Steady-state time per pass of the reference's loop shape through examples/sppm_render (4 replicas x (eye pass +
# 10000 x 100 photons) per pass, 1024x768), for each "NAME|ENV=.. ENV=.." argument: (12-pass total - 2-pass total) / 10
# of the host phase "pass (everything)" (SPPM_HOST_TIMING=1).
#   gpurun --timeout 600 -- 'bash tools/gpu_cpp_loop.sh r2_p "base|" "b256k|SPPM_PHOTON_BATCH=262144"'
set -u
TAG=$1; shift
OUT=$PWD/gpurun_out; mkdir -p $OUT
W=$(mktemp -d); cp examples/scenes/reference.scene $W/ref.scene
python -c "from raytrace3_b200.scene import teapot_patches, write_bpt; write_bpt('$W/teapot.bpt', teapot_patches())"
EXE=$PWD/examples/sppm_render
cd $W
for spec in "$@"; do
  name=${spec%%|*}; envs=${spec#*|}
  for P in 2 12; do
    env $envs SPPM_HOST_TIMING=1 $EXE ref.scene --size 1024x768 --passes $P --replicas 4 --rounds 10000 --photons 100 --newton fast --math fast --gpus 0 --out $W/o > $W/log_$P 2>&1 || { echo "$name FAILED"; tail -3 $W/log_$P; }
  done
  python - "$name" $W/log_2 $W/log_12 <<'PY' | tee -a $OUT/${TAG}_cpp_loop.txt
import re, sys
def phases(p):
    d = {}
    for l in open(p):
        m = re.match(r"host phase (.+?)\s+([\d.]+) ms total\s+(\d+) calls", l)
        if m: d[m.group(1).strip()] = (float(m.group(2)), int(m.group(3)))
    return d
a, b = phases(sys.argv[2]), phases(sys.argv[3])
row = []
for k in ("pass (everything)", "replica (everything)", "eye pass", "build index", "photon passes (enqueue)", "resolve + read-back"):
    if k in a and k in b:
        row.append(f"{k} {(b[k][0] - a[k][0]) / (b[k][1] - a[k][1]):.2f}")
print(f"{sys.argv[1]:10s} steady-state ms per call: " + " | ".join(row))
PY
done
