#!/usr/bin/env bash
# The two sharded BASELINE configs (4: caustics-1080p, 5: stress-4k PROGRESSIVE) at N GPUs, strong scaling.
#   gpurun --gpus 8 --timeout 900 -- 'bash tools/gpu_n8_sharded.sh r2_ZZ 8'
set -u
TAG=$1; N=$2
OUT=gpurun_out; mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
for spec in "caustics-1080p 10" "stress-4k 6"; do
  set -- $spec
  f=$OUT/${TAG}_$1_n$N.json
  timeout 400 $TR bench.py --gpus $N --workload $1 --steps $2 --warmup 4 --no-cpu-baseline --no-exact-mode --e2e-steps 2 > $f 2> $OUT/${TAG}_$1_n$N.err || { echo "$1 FAILED"; tail -5 $OUT/${TAG}_$1_n$N.err; continue; }
  python - "$f" <<'PY'
import json, sys
d = json.load(open(sys.argv[1]))
print(f"N={d['n_gpus']} {d['config']['workload']:16s} {d['scaling']:6s} value {d['value']/1e6:9.1f} M/s  ms/step {d['ms_per_step']:8.3f}  e2e {d['e2e']['value']/1e6:8.1f}  allreduce {d['allreduce_ms_per_step']}  parity {d['multi_gpu_parity'] if isinstance(d['multi_gpu_parity'], str) else (d['multi_gpu_parity'] or {}).get('status')}")
PY
done
