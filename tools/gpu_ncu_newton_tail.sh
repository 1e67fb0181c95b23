set -u
OUT=gpurun_out; mkdir -p $OUT
NAME=r2_O_newton_tail
python tools/ncu_target.py --region round > /dev/null 2>&1 || exit 1
timeout 300 ncu --set full --clock-control none --import-source on --profile-from-start off -f -k regex:newton_items -s 7 -c 1 -o $OUT/$NAME python tools/ncu_target.py --region round > $OUT/${NAME}_ncu.log 2>&1
python tools/ncu_key.py $OUT/$NAME.ncu-rep > $OUT/${NAME}_key.txt 2>&1
ncu -i $OUT/$NAME.ncu-rep --page source --csv --print-source cuda,sass > $OUT/${NAME}_src.csv 2>/dev/null
python tools/ncu_lines.py $OUT/${NAME}_src.csv 40 > $OUT/${NAME}_lines.txt 2>&1
ncu -i $OUT/$NAME.ncu-rep --page details --csv > $OUT/${NAME}_details.csv 2>/dev/null
rm -f $OUT/${NAME}_src.csv
cat $OUT/${NAME}_key.txt
