#!/usr/bin/env bash
# Round-end evidence on one GPU: GPU tests, bench + reference arm, launch lists, ncu --set full of the three
# largest kernel classes (raw pages exported on the box), the digest of the kernel sources they were taken from.
#   gpurun --timeout 1500 -- 'bash tools/gpu_final.sh r2_Z'
set -u
TAG=${1:-r2_Z}
OUT=gpurun_out; mkdir -p $OUT
T="python tools/ncu_target.py"
NCU="ncu --set full --clock-control none --import-source on --profile-from-start off -f"
timeout 700 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > $OUT/${TAG}_gpu_tests.txt; cat $OUT/${TAG}_gpu_tests.txt
for w in cornell-1024x768 caustics-1080p stress-4k ref-main-320x240; do
  timeout 400 python bench.py --workload $w --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 1 > $OUT/${TAG}_bench_$w.json 2> $OUT/${TAG}_err_$w.txt
  cut -c1-160 $OUT/${TAG}_bench_$w.json
done
$T --region round > /dev/null 2>&1 || { echo plain failed; exit 1; }
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $OUT/${TAG}_launches_round.csv $T --region round > /dev/null 2>&1
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $OUT/${TAG}_launches_build.csv $T --region build > /dev/null 2>&1
cap() {  # name kernel-regex skip count
  timeout 300 $NCU -k regex:$2 -s $3 -c $4 -o $OUT/${TAG}_$1 $T --region round > $OUT/${TAG}_ncu_$1.log 2>&1
  ncu -i $OUT/${TAG}_$1.ncu-rep --page raw --csv > $OUT/${TAG}_$1_raw.csv 2>/dev/null
  python tools/ncu_key.py $OUT/${TAG}_$1.ncu-rep > $OUT/${TAG}_$1_key.txt 2>&1
  ncu -i $OUT/${TAG}_$1.ncu-rep --page source --csv --print-source cuda,sass > $OUT/${TAG}_$1_src.csv 2>/dev/null
  python tools/ncu_lines.py $OUT/${TAG}_$1_src.csv 45 > $OUT/${TAG}_$1_lines.txt 2>&1
  rm -f $OUT/${TAG}_$1_src.csv $OUT/${TAG}_$1.ncu-rep
}
cap gather gather_filtered 0 1
cap step photon_step 0 15
cap newton newton_items 0 14
python -c "import bench; print(bench.csrc_digest())" > $OUT/${TAG}_csrc_digest.txt
# the traffic table of THIS kernel state, then the bench line that quotes it
python tools/ncu_traffic.py $(cat $OUT/${TAG}_csrc_digest.txt) teapot-room-1024x768 $OUT/${TAG}_gather_raw.csv $OUT/${TAG}_step_raw.csv $OUT/${TAG}_newton_raw.csv > /dev/null && cp profiles/ncu_traffic.json $OUT/${TAG}_ncu_traffic.json
timeout 400 python bench.py --steps 20 --warmup 5 > $OUT/${TAG}_bench_n1.json 2> $OUT/${TAG}_bench_n1.err; cut -c1-200 $OUT/${TAG}_bench_n1.json
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/${TAG}_bench_reference_arm.json 2>/dev/null; cut -c1-200 $OUT/${TAG}_bench_reference_arm.json
du -sh $OUT
