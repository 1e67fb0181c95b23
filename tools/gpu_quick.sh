#!/usr/bin/env bash
# Quick GPU check: GPU tests + one bench line (+ optional extra bench arguments as $2..).
#   gpurun --timeout 900 -- 'bash tools/gpu_quick.sh r2_C'
set -u
TAG=${1:-r2_X}; shift || true
OUT=gpurun_out
mkdir -p $OUT
timeout 700 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > $OUT/${TAG}_gpu_tests.txt
cat $OUT/${TAG}_gpu_tests.txt
timeout 400 python bench.py --steps 20 --warmup 5 "$@" > $OUT/${TAG}_bench_n1.json 2> $OUT/${TAG}_bench_n1.err
echo "bench rc=$?"; cut -c1-400 $OUT/${TAG}_bench_n1.json; tail -3 $OUT/${TAG}_bench_n1.err
python -c "import bench; print(bench.csrc_digest())" > $OUT/${TAG}_csrc_digest.txt
