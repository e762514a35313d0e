#!/bin/bash
# usage: scripts/ncu_quick.sh name [name ...] — headline ncu metrics of k_render for ab/<name>.so (tuning aid; 1000x1000x4 spp)
M=smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__inst_executed.sum,gpu__time_duration.sum,sm__warps_active.avg.pct_of_peak_sustained_active,l1tex__t_sector_hit_rate.pct,smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio,smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio,smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_membar_per_issue_active.ratio,smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio,launch__registers_per_thread
for n in "$@"; do
  echo "== $n"
  RTNW_LIB=ab/$n.so timeout 300 ncu --metrics $M --clock-control none -k regex:k_render -c 1 python scripts/prof_render.py --ns 4 --reps 1 2>&1 | grep -E "^\s+(smsp|gpu__|sm__|l1tex|launch)" | sed 's/  */ /g'
done
