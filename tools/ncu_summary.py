"""Selected metrics of one `ncu --set full` capture -> small CSV (what profiles/ keeps).
usage: ncu_summary.py <report.ncu-rep> <out.csv>"""
import csv
import subprocess
import sys

KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__shared_mem_per_block_dynamic", "lts__t_sector_hit_rate.pct",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__cycles_elapsed.max", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "smsp__cycles_active.avg", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum.per_second", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_subpipe_imma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_subpipe_imma_cycles_active_realtime.avg",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu_realtime.avg.pct_of_peak_sustained_elapsed"]
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
out = [("Kernel Name", "", vals[hdr.index("Kernel Name")])]
for i, h in enumerate(hdr):
    base = h.split(".TriageCompute.")[-1]
    if base in KEEP:
        out.append((h, units[i], vals[i]))
with open(sys.argv[2], "w") as f:
    w = csv.writer(f)
    w.writerow(["metric", "unit", "value (one warm launch, ncu --set full --clock-control none)"])
    w.writerows(out)
for r in out:
    print(*r)
