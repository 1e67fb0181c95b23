"""Summarise `ncu --page source --csv --print-source cuda,sass` output per source line:
instructions executed, warp-stall samples, share.  Usage: ncu_lines.py file.csv [top_n]"""
import csv
import sys
from collections import defaultdict

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur = None
per_file = defaultdict(lambda: [0, 0])
lines = []
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if len(r) > 7 and r[0].isdigit():
        try:
            inst = int(r[7])
            samp = int(r[4]) if r[4].isdigit() else 0
        except ValueError:
            continue
        lines.append((inst, samp, cur, int(r[0]), r[1].strip()[:110]))
        per_file[cur][0] += inst
        per_file[cur][1] += samp
tot_i = sum(x[0] for x in lines) or 1
tot_s = sum(x[1] for x in lines) or 1
print(f"total warp-instructions {tot_i:,}  stall samples {tot_s:,}")
for f, (i, s) in sorted(per_file.items(), key=lambda kv: -kv[1][0]):
    print(f"  {f:24s} inst {i / tot_i:6.1%}  samples {s / tot_s:6.1%}")
print("-- by instructions")
for inst, samp, f, ln, src in sorted(lines, key=lambda x: -x[0])[:top]:
    print(f"{inst / tot_i:6.1%} i {samp / tot_s:6.1%} s  {f}:{ln:<4d} {src}")
print("-- by stall samples")
for inst, samp, f, ln, src in sorted(lines, key=lambda x: -x[1])[:top]:
    print(f"{samp / tot_s:6.1%} s {inst / tot_i:6.1%} i  {f}:{ln:<4d} {src}")
