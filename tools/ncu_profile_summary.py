"""Write the committed ncu summary of one `ncu --set full --import-source on` capture (profiles/*_ncu_full.txt): launch
configuration, issue / cache / DRAM counters, warp-stall breakdown (cycles per issued instruction) and the attribution of
stall samples / executed instructions to source functions.
Usage: ncu_profile_summary.py capture.ncu-rep "command line that was profiled" out.txt"""
import csv, io, os, subprocess, sys
rep, cmd, out_path = sys.argv[1], sys.argv[2], sys.argv[3]
HERE = os.path.dirname(os.path.abspath(__file__))
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h, u, r = rows[0], rows[1], rows[2]
want = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps",
        "sm__warps_active.avg.per_cycle_active", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed.avg.per_cycle_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum", "l1tex__t_sector_hit_rate.pct",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld_lookup_hit.sum", "lts__t_sector_hit_rate.pct",
        "smsp__sass_inst_executed_op_global_ld.sum", "smsp__sass_inst_executed_op_global_st.sum", "smsp__sass_inst_executed_op_local_ld.sum",
        "smsp__sass_inst_executed_op_local_st.sum", "smsp__sass_inst_executed_op_shared_ld.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__icc_request_hit_rate.pct", "gcc__cache_requests_type_instruction.sum.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tc.avg.pct_of_peak_sustained_active"]
out = ["ncu --set full --import-source on --clock-control none -k regex:k_step -s 1 -c 1 " + cmd, "kernel: " + r[h.index("Kernel Name")], ""]
for w in want:
    if w in h:
        i = h.index(w); out.append("  %-82s %-16s %s" % (w, u[i], r[i]))
out += ["", "  warp stall reasons, cycles per issued instruction (smsp__average_warps_issue_stalled_*_per_issue_active.ratio):"]
st = [(float(r[i]), x[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]) for i, x in enumerate(h)
      if x.startswith("smsp__average_warps_issue_stalled_") and x.endswith("_per_issue_active.ratio")]
tot = sum(v for v, _ in st)
out += ["    %-22s %6.2f  (%4.1f%%)" % (n, v, 100 * v / tot) for v, n in sorted(st, reverse=True) if v >= 0.01]
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
tmp = out_path + ".src.csv"
open(tmp, "w").write(src)
f = subprocess.run([sys.executable, os.path.join(HERE, "ncu_funcs.py"), tmp, "22"], capture_output=True, text=True).stdout
os.remove(tmp)
out += ["", "  source attribution (innermost inlined function; share of stall samples / of executed warp instructions):"] + ["  " + l for l in f.splitlines()]
open(out_path, "w").write("\n".join(out) + "\n")
print(out_path, "time", r[h.index("gpu__time_duration.sum")], "ms; dram r/w", r[h.index("dram__bytes_read.sum")], u[h.index("dram__bytes_read.sum")],
      r[h.index("dram__bytes_write.sum")], u[h.index("dram__bytes_write.sum")], "; inst", r[h.index("smsp__inst_executed.sum")])
