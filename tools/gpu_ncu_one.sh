#!/usr/bin/env bash
# One ncu --set full capture (with source) of one kernel in one region of tools/ncu_target.py.
#   gpurun --timeout 600 -- 'bash tools/gpu_ncu_one.sh r2_H_gather4k gather_filtered "--workload stress-4k --photons 2000000"'
set -u
NAME=$1; KERN=$2; ARGS=${3:-}
OUT=gpurun_out; mkdir -p $OUT
python tools/ncu_target.py --region round $ARGS > $OUT/${NAME}_plain.log 2>&1 || { echo plain failed; tail -5 $OUT/${NAME}_plain.log; exit 1; }
timeout 400 ncu --set full --clock-control none --import-source on --profile-from-start off -f -k regex:$KERN -c ${4:-1} -o $OUT/$NAME python tools/ncu_target.py --region round $ARGS > $OUT/${NAME}_ncu.log 2>&1
ncu -i $OUT/$NAME.ncu-rep --page raw --csv > $OUT/${NAME}_raw.csv 2>/dev/null
python tools/ncu_key.py $OUT/$NAME.ncu-rep > $OUT/${NAME}_key.txt 2>&1
ncu -i $OUT/$NAME.ncu-rep --page source --csv --print-source cuda,sass > $OUT/${NAME}_src.csv 2>/dev/null
python tools/ncu_lines.py $OUT/${NAME}_src.csv 60 > $OUT/${NAME}_lines.txt 2>&1
rm -f $OUT/${NAME}_src.csv; [ $(stat -c %s $OUT/$NAME.ncu-rep) -gt 9000000 ] && rm -f $OUT/$NAME.ncu-rep
python -c "import bench; print(bench.csrc_digest())" > $OUT/${NAME}_csrc_digest.txt
cat $OUT/${NAME}_key.txt | head -30
