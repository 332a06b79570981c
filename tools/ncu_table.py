"""One-line-per-report table of the metrics DESIGN.md / profiles/ quote."""
import csv, subprocess, sys, io
keys = [("gpu__time_duration.sum", "ms"), ("dram__bytes_read.sum", "GB rd"), ("dram__bytes_write.sum", "GB wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
        ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem wavefronts"),
        ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "bank conflicts"),
        ("l1tex__throughput.avg.pct_of_peak_sustained_active", "l1tex% (active)"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
        ("smsp__inst_executed.sum", "warp insts"),
        ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long_sb"),
        ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall short_sb"),
        ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall barrier"),
        ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "stall mio"),
        ("launch__registers_per_thread", "regs"), ("launch__shared_mem_per_block_dynamic", "dyn smem KB"),
        ("launch__grid_size", "grid")]
for rep in sys.argv[1:]:
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, vals = rows[0], rows[2]
    print("##", rep, "--", vals[hdr.index("Kernel Name")][:48])
    for k, label in keys:
        if k in hdr:
            print(f"  {label:18s} {vals[hdr.index(k)]:>16s} {rows[1][hdr.index(k)]}")
