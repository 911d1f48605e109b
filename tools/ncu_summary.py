"""Summarise an ncu report (--set full) into a small CSV for profiles/: one row per profiled launch with the metrics the
roofline discussion uses.  Usage: ncu_summary.py report.ncu-rep out.csv"""
import csv, subprocess, sys
METRICS = ["gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
           "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__inst_executed_pipe_uniform.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "smsp__cycles_active.avg",
           "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
           "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
           "launch__shared_mem_per_block_dynamic", "sm__cycles_active.avg", "smsp__inst_executed.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
with open(sys.argv[2], "w", newline="") as f:
    w = csv.writer(f)
    cols = [m for m in METRICS if m in idx]
    w.writerow(["kernel"] + [f"{m} [{units[idx[m]]}]" for m in cols])
    for r in rows[2:]:
        w.writerow([r[idx["Kernel Name"]]] + [r[idx[m]] for m in cols])
print("wrote", sys.argv[2], len(rows) - 2, "launches")
