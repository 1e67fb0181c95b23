#!/usr/bin/env bash
# Round-2 GPU session B: GPU tests (incl. the bench-mode parity tests), bench line, serialised launch lists,
# ncu --set full of newton_items / patch_batch / the build_index chain / gather / step kernels.  The raw pages
# are exported to CSV ON THE BOX and reports above 9 MB are dropped (gpurun_out/ is capped at 64 MiB).
#   gpurun --timeout 1500 -- 'bash tools/gpu_r2b.sh r2_B'
set -u
TAG=${1:-r2_B}
WHAT=${2:-all}
OUT=gpurun_out
mkdir -p $OUT
NCU="ncu --set full --clock-control none --profile-from-start off -f"
T="python tools/ncu_target.py"
want() { [ "$WHAT" = all ] || [ "$WHAT" = "$1" ]; }
export_rep() {  # raw page -> csv + key metrics; source page for reports captured with --import-source on
  local rep=$OUT/$1.ncu-rep
  [ -f $rep ] || { echo "no $rep"; return; }
  ncu -i $rep --page raw --csv > $OUT/$1_raw.csv 2>/dev/null
  python tools/ncu_key.py $rep > $OUT/$1_key.txt 2>&1
  if [ "${2:-}" = src ]; then
    ncu -i $rep --page source --csv --print-source cuda,sass > $OUT/$1_src.csv 2>/dev/null
    python tools/ncu_lines.py $OUT/$1_src.csv 45 > $OUT/$1_lines.txt 2>&1
    [ $(stat -c %s $OUT/$1_src.csv) -gt 6000000 ] && rm -f $OUT/$1_src.csv
  fi
  [ $(stat -c %s $rep) -gt 9000000 ] && rm -f $rep
  [ $(stat -c %s $OUT/$1_raw.csv) -gt 4000000 ] && gzip -f $OUT/$1_raw.csv
}

if want tests; then
  timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > $OUT/${TAG}_gpu_tests.txt
  cat $OUT/${TAG}_gpu_tests.txt
fi
if want bench; then
  timeout 300 python bench.py --steps 20 --warmup 5 > $OUT/${TAG}_bench_n1.json 2> $OUT/${TAG}_bench_n1.err
  cut -c1-300 $OUT/${TAG}_bench_n1.json; tail -2 $OUT/${TAG}_bench_n1.err
fi
if want ncu; then
  # plain runs first (a region must exit 0 without ncu before it is profiled)
  $T --region round > $OUT/${TAG}_plain_round.log 2>&1 && $T --region build > $OUT/${TAG}_plain_build.log 2>&1 && $T --region patch > $OUT/${TAG}_plain_patch.log 2>&1 || { echo "plain target failed"; tail -5 $OUT/${TAG}_plain_*.log; exit 1; }
  timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $OUT/${TAG}_launches_round.csv $T --region round > $OUT/${TAG}_ncu_l1.log 2>&1
  timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $OUT/${TAG}_launches_build.csv $T --region build > $OUT/${TAG}_ncu_l2.log 2>&1
  timeout 300 $NCU -k regex:newton_items -o $OUT/${TAG}_newton $T --region round > $OUT/${TAG}_ncu_newton.log 2>&1
  export_rep ${TAG}_newton
  timeout 200 $NCU --import-source on -k regex:newton_items -c 1 -o $OUT/${TAG}_newton1 $T --region round > $OUT/${TAG}_ncu_newton1.log 2>&1
  export_rep ${TAG}_newton1 src
  timeout 200 $NCU --import-source on -k regex:patch_batch -c 1 -o $OUT/${TAG}_patch $T --region patch > $OUT/${TAG}_ncu_patch.log 2>&1
  export_rep ${TAG}_patch src
  timeout 400 $NCU -o $OUT/${TAG}_build $T --region build > $OUT/${TAG}_ncu_build.log 2>&1
  export_rep ${TAG}_build
  timeout 200 $NCU --import-source on -k regex:gather_filtered -c 1 -o $OUT/${TAG}_gather $T --region round > $OUT/${TAG}_ncu_gather.log 2>&1
  export_rep ${TAG}_gather src
  timeout 300 $NCU -k regex:photon_step -o $OUT/${TAG}_step $T --region round > $OUT/${TAG}_ncu_step.log 2>&1
  export_rep ${TAG}_step
  timeout 200 $NCU --import-source on -k regex:photon_step -s 1 -c 1 -o $OUT/${TAG}_step1 $T --region round > $OUT/${TAG}_ncu_step1.log 2>&1
  export_rep ${TAG}_step1 src
fi
python -c "import bench; print(bench.csrc_digest())" > $OUT/${TAG}_csrc_digest.txt
rm -f $OUT/*.ncu-rep.tmp
du -sh $OUT; ls -la $OUT | grep ${TAG} | awk '{print $5, $9}'
