#!/bin/bash
# quick counter set for one stepping launch: tools/ncu_quick.sh <tag> <cfg> <worlds>   (env: B2G_LPW, B2G_LANES, B2G_LIB)
M=gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__icc_request_hit_rate.pct,gcc__cache_requests_type_instruction.sum.pct_of_peak_sustained_elapsed,launch__registers_per_thread,sm__warps_active.avg.per_cycle_active,l1tex__t_sector_hit_rate.pct,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio,smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio,smsp__average_warps_issue_stalled_membar_per_issue_active.ratio,smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio,smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio,smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio,smsp__average_warps_issue_stalled_misc_per_issue_active.ratio
ncu --metrics $M --clock-control none -k regex:k_step -s 1 -c 1 --csv --log-file gpurun_out/q_$1.csv python tools/prof_cfg.py $2 $3 2 > gpurun_out/q_$1.log 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open("gpurun_out/q_$1.csv")) if len(r)>10]
h=rows[0]; 
print("== $1")
for r in rows[1:]:
    n=r[h.index("Metric Name")]; v=r[h.index("Metric Value")]
    n=n.replace("smsp__average_warps_issue_stalled_","stall ").replace("_per_issue_active.ratio","")
    try:
        if float(v.replace(",",""))>=0.05: print("  %-70s %s"%(n,v))
    except: print("  %-70s %s"%(n,v))
PY
