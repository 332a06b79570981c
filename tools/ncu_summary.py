"""Summarise an ncu --set full report (raw page) into the handful of metrics we judge by.
usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep [more-metric-substrings...]"""
import csv, subprocess, sys, io
rep = sys.argv[1]
extra = sys.argv[2:]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared", "smsp__inst_executed.sum", "sm__warps_active.avg.pct",
        "launch__registers_per_thread", "sm__cycles_elapsed.avg", "sm__cycles_active.avg", "lts__t_bytes.sum",
        "smsp__average_warp", "smsp__issue_active.avg.pct", "l1tex__data_pipe_lsu_wavefronts.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum", "smsp__inst_executed_op_shared_ld",
        "l1tex__throughput.avg.pct", "lts__throughput.avg.pct", "launch__shared_mem", "launch__grid_size",
        "smsp__warps_issue_stalled", "sm__inst_executed_pipe_lsu.sum", "l1tex__lsu_writeback_active"] + extra
for vals in rows[2:]:
    print("==", vals[hdr.index("Kernel Name")][:60], "grid", vals[hdr.index("Grid Size")], "block", vals[hdr.index("Block Size")])
    for h, u, v in zip(hdr, units, vals):
        if any(w in h for w in want):
            print(f"  {h} [{u}] = {v}")
