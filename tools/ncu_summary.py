"""Print the metrics that matter for this workload from an `ncu --page raw --csv` dump."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "launch__occupancy_limit", "launch__grid_size", "launch__block_size",
        "smsp__issue_active.avg.pct", "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__warp_issue_stalled",
        "l1tex__data_bank_conflicts_pipe_lsu", "lts__throughput.avg.pct", "l1tex__throughput.avg.pct",
        "launch__shared_mem_per_block", "smsp__inst_executed_op_shared", "smsp__inst_executed_op_global", "smsp__inst_executed_op_local",
        "dram__throughput", "sm__cycles_elapsed.max", "l1tex__t_bytes", "smsp__sass_thread_inst_executed_op_fp32",
        "sm__sass_inst_executed_op_shared", "derived__smsp__sass_thread_inst_executed_op_ffma", "sm__inst_executed_pipe"]
for r in rows[2:]:
    print("==== kernel:", r[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "")
    for i, h in enumerate(hdr):
        if any(w in h for w in want):
            if "stalled" in h and "per_warp_active" not in h:
                continue
            print("  %-90s %-14s %s" % (h, units[i], r[i]))
