#!/usr/bin/env bash
# Multi-GPU session (gpurun --gpus N): correctness on hardware, then bench lines at N GPUs.
#   gpurun --gpus 2 --timeout 900 -- 'bash tools/gpu_multi.sh r2_E 2 "base||" "sync|SPPM_ASYNC_COMM=0|"'
set -u
TAG=$1; N=$2; shift 2
OUT=gpurun_out
mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 300 $TR tools/check_multigpu.py > $OUT/${TAG}_check_multigpu_n$N.txt 2>&1
echo "check_multigpu rc=$?"; tail -6 $OUT/${TAG}_check_multigpu_n$N.txt
for spec in "$@"; do
  name=${spec%%|*}; rest=${spec#*|}; envs=${rest%%|*}; args=${rest#*|}
  f=$OUT/${TAG}_${name}_n$N.json
  env $envs timeout 500 $TR bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline $args > $f 2> $OUT/${TAG}_${name}_n$N.err || { echo "$name FAILED"; tail -5 $OUT/${TAG}_${name}_n$N.err; continue; }
  python - "$name" "$f" <<'PY'
import json, sys
d = json.load(open(sys.argv[2]))
k = d["kernels"]
print(f"{sys.argv[1]:12s} N={d['n_gpus']} {d['config']['workload']:22s} {d['scaling']:6s} value {d['value']/1e6:9.1f} M/s  ms/step {d['ms_per_step']:8.3f}  e2e {d['e2e']['value']/1e6:8.1f}  allreduce {d['allreduce_ms_per_step']}  parity {d['multi_gpu_parity'] if isinstance(d['multi_gpu_parity'], str) else (d['multi_gpu_parity'] or {}).get('status')}")
PY
done
