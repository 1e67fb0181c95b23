#!/usr/bin/env bash
# Round-2 first GPU session: GPU tests, compute-sanitizer, ncu --set full of the kernels
# round 1 left without counters (newton_items, patch_batch, the build_index chain), a serialised launch
# list, and a baseline bench line.   gpurun --timeout 1500 -- 'bash tools/gpu_r2a.sh r2_A'
set -u
TAG=${1:-r2_A}
OUT=gpurun_out
mkdir -p $OUT
NCU="ncu --set full --clock-control none --import-source on --profile-from-start off -f"
T="python tools/ncu_target.py"

timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > $OUT/${TAG}_gpu_tests.txt
cat $OUT/${TAG}_gpu_tests.txt

timeout 500 compute-sanitizer --tool memcheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_memcheck_smoke.log 2>&1
echo "memcheck smoke rc=$?"; tail -3 $OUT/${TAG}_memcheck_smoke.log
timeout 500 compute-sanitizer --tool racecheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_racecheck_smoke.log 2>&1
echo "racecheck smoke rc=$?"; tail -3 $OUT/${TAG}_racecheck_smoke.log
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -x -q -k "gather_filtered_fuzz or gather_edge_cases" > $OUT/${TAG}_memcheck_gather.log 2>&1
echo "memcheck gather rc=$?"; tail -3 $OUT/${TAG}_memcheck_gather.log
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -x -q -k "gather_filtered_fuzz and 0" > $OUT/${TAG}_racecheck_gather.log 2>&1
echo "racecheck gather rc=$?"; tail -3 $OUT/${TAG}_racecheck_gather.log

timeout 240 python bench.py --steps 20 --warmup 5 > $OUT/${TAG}_bench_n1.json 2> $OUT/${TAG}_bench_n1.err
cut -c1-400 $OUT/${TAG}_bench_n1.json

# plain runs first (a region must exit 0 without ncu before it is profiled)
$T --region round > $OUT/${TAG}_plain_round.log 2>&1 && $T --region build > $OUT/${TAG}_plain_build.log 2>&1 && $T --region patch > $OUT/${TAG}_plain_patch.log 2>&1 || { echo "plain target failed"; tail -5 $OUT/${TAG}_plain_*.log; exit 1; }

timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $OUT/${TAG}_launches_round.csv $T --region round > $OUT/${TAG}_ncu_l1.log 2>&1
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $OUT/${TAG}_launches_build.csv $T --region build > $OUT/${TAG}_ncu_l2.log 2>&1
timeout 300 $NCU -k regex:newton_items -o $OUT/${TAG}_newton $T --region round > $OUT/${TAG}_ncu_newton.log 2>&1
timeout 200 $NCU -k regex:patch_batch -c 1 -o $OUT/${TAG}_patch $T --region patch > $OUT/${TAG}_ncu_patch.log 2>&1
timeout 400 $NCU -o $OUT/${TAG}_build $T --region build > $OUT/${TAG}_ncu_build.log 2>&1
timeout 200 $NCU -k regex:gather_filtered -c 1 -o $OUT/${TAG}_gather $T --region round > $OUT/${TAG}_ncu_gather.log 2>&1
timeout 300 $NCU -k regex:photon_step -c 4 -o $OUT/${TAG}_step $T --region round > $OUT/${TAG}_ncu_step.log 2>&1
ls -la $OUT/${TAG}_*.ncu-rep
