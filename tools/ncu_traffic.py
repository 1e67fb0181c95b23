#!/usr/bin/env python
"""profiles/ncu_traffic.json from `ncu --page raw --csv` exports: dram__bytes_read.sum + dram__bytes_write.sum per
launch of each kernel, with the digest of the kernel sources they were captured from (bench.py quotes a number
only while the sources still match).

  python tools/ncu_traffic.py <csrc_digest> <workload> gpurun_out/r2_B_gather_raw.csv [more.csv ...]
"""
import csv
import gzip
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def rows_of(path):
    op = gzip.open if path.endswith(".gz") else open
    with op(path, "rt") as f:
        rows = list(csv.reader(f))
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        yield dict(zip(hdr, vals)), dict(zip(hdr, units))


def main():
    digest, workload, files = sys.argv[1], sys.argv[2], sys.argv[3:]
    out_path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    out = {"csrc_digest": digest, "kernels": {}}
    if os.path.exists(out_path):
        old = json.load(open(out_path))
        if old.get("csrc_digest") == digest:
            out = old
    acc = {}
    for path in files:
        for d, u in rows_of(path):
            name = re.sub(r"<.*", "", d.get("Kernel Name", "").split("(")[0]).split("::")[-1]
            tot = 0.0
            for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                tot += float(d[k].replace(",", "")) * UNIT.get(u[k], 1.0)
            t = float(d["gpu__time_duration.sum"].replace(",", "")) * {"us": 1e-3, "ms": 1.0, "ns": 1e-6, "usecond": 1e-3, "msecond": 1.0, "nsecond": 1e-6}.get(u["gpu__time_duration.sum"], 1e-3)
            a = acc.setdefault(name, dict(n=0, bytes=0.0, ms=0.0, src=os.path.basename(path)))
            a["n"] += 1
            a["bytes"] += tot
            a["ms"] += t
    for name, a in acc.items():
        out["kernels"][f"{name}@{workload}"] = {"dram_bytes": a["bytes"] / a["n"], "launches_captured": a["n"],
                                                 "ms_under_ncu": a["ms"] / a["n"], "source": f"profiles/{a['src']} (ncu --set full, one launch each)"}
    json.dump(out, open(out_path, "w"), indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
