#!/usr/bin/env bash
# A/B runs of bench.py on one GPU: each argument is "NAME|ENV=.. ENV=..|bench args"; prints value, ms/step
# and the serialised per-kernel times.   gpurun --timeout 900 -- 'bash tools/gpu_ab.sh r2_D "base||" "tail4|SPPM_WF_TAIL=4|"'
set -u
TAG=$1; shift
OUT=gpurun_out
mkdir -p $OUT
for spec in "$@"; do
  name=${spec%%|*}; rest=${spec#*|}; envs=${rest%%|*}; args=${rest#*|}
  f=$OUT/${TAG}_${name}.json
  env $envs timeout 400 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --e2e-steps 0 $args > $f 2> $OUT/${TAG}_${name}.err || { echo "$name FAILED"; tail -3 $OUT/${TAG}_${name}.err; continue; }
  python - "$name" "$f" <<'PY'
import json, sys
d = json.load(open(sys.argv[2]))
k = d["kernels"]
print(f"{sys.argv[1]:14s} {d['config']['workload']:22s} value {d['value']/1e6:8.1f} M/s  ms/step {d['ms_per_step']:8.3f}  exact {(d['exact_mode'] or {}).get('value', 0)/1e6:7.1f}  ser_round {d['roofline']['serialized_round_ms']:7.3f} | " +
      "  ".join(f"{n.split('_kernel')[0].split(' ')[0]} {v['ms_per_step']:.3f}" for n, v in k.items()))
PY
done
