"""Selected counters of the first kernel of an `ncu --set full` report as `metric,unit,value` rows (the format of profiles/ncu_*.csv).
    python tests/tools/ncu_counters.py report.ncu-rep > profiles/ncu_rNN_<kernel>.csv"""
import csv, subprocess, sys
WANT = ["dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "launch__block_size", "launch__grid_size", "launch__cluster_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "sass__inst_executed_local_loads", "sass__inst_executed_local_stores", "sm__cycles_elapsed.max",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "lts__t_bytes.sum", "l1tex__m_xbar2l1tex_read_bytes.sum"]
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
w = csv.writer(sys.stdout); w.writerow(["metric", "unit", "value"])
w.writerow(["Kernel Name", "", vals[hdr.index("Kernel Name")]])
for i, h in enumerate(hdr):
    if h in WANT or h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio"):
        w.writerow([h, units[i], vals[i]])
