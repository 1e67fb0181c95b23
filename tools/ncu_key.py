"""Print the handful of ncu raw-page metrics the design notes quote.  Usage: ncu_key.py report.ncu-rep"""
import csv
import io
import subprocess
import sys

WANT = """gpu__time_duration.sum dram__bytes_read.sum dram__bytes_write.sum sm__throughput.avg.pct_of_peak_sustained_elapsed
sm__warps_active.avg.pct_of_peak_sustained_active launch__registers_per_thread launch__grid_size launch__block_size
sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active
sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active
smsp__inst_executed.sum smsp__thread_inst_executed_per_inst_executed.ratio smsp__issue_active.avg.pct_of_peak_sustained_active
smsp__warps_eligible.avg.per_cycle_active l1tex__data_pipe_lsu_wavefronts_mem_shared.sum l1tex__throughput.avg.pct_of_peak_sustained_active
lts__t_sector_hit_rate.pct lts__throughput.avg.pct_of_peak_sustained_elapsed gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed
l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum sm__cycles_elapsed.avg""".split()
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
for vals in rows[2:]:
    d = dict(zip(hdr, vals))
    print("##", d.get("Kernel Name", "")[:90])
    for h, u in zip(hdr, units):
        if h in WANT:
            print(f"  {h:75s} {d[h]:>16s} {u}")
    st = sorted(((float(d[h]), h.split("issue_stalled_")[1].split("_per_")[0]) for h in hdr
                 if "issue_stalled" in h and h.endswith("per_issue_active.ratio") and d[h]), reverse=True)
    print("  stalls per issue:", ", ".join(f"{n} {v:.2f}" for v, n in st[:8]))
