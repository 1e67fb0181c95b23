#!/usr/bin/env bash
# A/B of the end-to-end leg (RayTracer.render()) on one GPU: each argument is "NAME|ENV=.. ENV=..|bench args".
#   gpurun --timeout 900 -- 'bash tools/gpu_e2e_ab.sh r2_k "base||" "b5m|SPPM_PHOTON_BATCH=5242880|"'
set -u
TAG=$1; shift
OUT=gpurun_out; mkdir -p $OUT
for spec in "$@"; do
  name=${spec%%|*}; rest=${spec#*|}; envs=${rest%%|*}; args=${rest#*|}
  f=$OUT/${TAG}_${name}.json
  env $envs timeout 400 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-exact-mode --e2e-steps 5 $args > $f 2> $OUT/${TAG}_${name}.err || { echo "$name FAILED"; tail -3 $OUT/${TAG}_${name}.err; continue; }
  python - "$name" "$f" <<'PY'
import json, sys
d = json.load(open(sys.argv[2])); e = d["e2e"]
print(f"{sys.argv[1]:10s} {d['config']['workload']:22s} e2e {e['value']/1e6:8.1f} M/s  {e['ms_per_step']:8.3f} ms per render  phases {e['phases_ms']}")
PY
done
