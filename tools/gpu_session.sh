#!/usr/bin/env bash
# One GPU-box session that produces everything a round's evidence needs, in the order that matters
# if the box time runs out.  Run it through gpurun from the repo root, e.g.
#   gpurun --timeout 900 -- 'bash tools/gpu_session.sh r2_A all'
# Stages (second argument, default "all"): tests | bench | launches | ncu | sanitize | configs.
# Outputs land in gpurun_out/<tag>_*; copy what should be judged into profiles/ afterwards
# (tools/ncu_key.py and tools/ncu_lines.py summarise the .ncu-rep files here, not on the box).
set -u
TAG=${1:-rX}
WHAT=${2:-all}
OUT=gpurun_out
mkdir -p $OUT
B="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --e2e-steps 0"
want() { [ "$WHAT" = all ] || [ "$WHAT" = "$1" ]; }

if want tests; then
  timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > $OUT/${TAG}_gpu_tests.txt
  cat $OUT/${TAG}_gpu_tests.txt
fi
if want bench; then
  timeout 240 python bench.py > $OUT/${TAG}_bench_n1.json 2> $OUT/${TAG}_bench_n1.err
  cut -c1-600 $OUT/${TAG}_bench_n1.json; tail -2 $OUT/${TAG}_bench_n1.err
  timeout 240 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/${TAG}_bench_reference_arm.json 2>&1
  cut -c1-400 $OUT/${TAG}_bench_reference_arm.json
fi
if want launches; then
  # launch list: only after the same command exited 0 without ncu; a number printed under ncu is never a bench value
  $B > $OUT/${TAG}_plain.log 2>&1 && timeout 240 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
      --log-file $OUT/${TAG}_launches.csv $B > $OUT/${TAG}_ncu_launches.log 2>&1
  tail -1 $OUT/${TAG}_ncu_launches.log | cut -c1-200
fi
if want ncu; then
  $B > $OUT/${TAG}_plain.log 2>&1 || exit 1
  # -s skips the warm-up rounds' launches of the same kernel (4 rounds x 15 bounces + the first of the timed round)
  timeout 120 ncu --set full --clock-control none --import-source on -k regex:gather_filtered -s 3 -c 1 -f -o $OUT/${TAG}_gather $B > $OUT/${TAG}_ncu_gather.log 2>&1
  timeout 120 ncu --set full --clock-control none --import-source on -k regex:photon_step -s 61 -c 2 -f -o $OUT/${TAG}_step $B > $OUT/${TAG}_ncu_step.log 2>&1
  timeout 120 ncu --set full --clock-control none --import-source on -k regex:newton_items -s 57 -c 2 -f -o $OUT/${TAG}_newton $B > $OUT/${TAG}_ncu_newton.log 2>&1
  ls -la $OUT/${TAG}_*.ncu-rep
fi
if want sanitize; then
  # memcheck over one small pass of every kernel on the path (smoke: eye pass, index build, photon pass, gather, update, resolve)
  timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_memcheck.log 2>&1
  echo "memcheck rc=$?"; tail -4 $OUT/${TAG}_memcheck.log
  timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_racecheck.log 2>&1
  echo "racecheck rc=$?"; tail -4 $OUT/${TAG}_racecheck.log
fi
if want configs; then
  for w in cornell-1024x768 caustics-1080p stress-4k ref-main-320x240; do
    timeout 240 python bench.py --workload $w --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 1 > $OUT/${TAG}_bench_$w.json 2> $OUT/${TAG}_err_$w.txt
    cut -c1-300 $OUT/${TAG}_bench_$w.json
  done
fi
