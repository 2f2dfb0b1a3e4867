#!/bin/bash
# ncu counters of the fused pass kernel with the gather role's phases switched off one at a time (probe build: -DFDQL_PROBES)
# usage: profiles/probe_ncu_roles.sh <lib> "<dbg values>"
lib=$1; shift
M=gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio,smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio,smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio,smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio,sm__icc_request_hit_rate.pct,sm__icc_requests.sum,l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,dram__bytes_read.sum,dram__bytes_write.sum
for d in $1; do
  FDQL_LIB=$lib FDQL_PROBE_DBG=$d ncu --metrics $M --clock-control none -k regex:fused_pass --launch-skip 40 -c 2 --csv --log-file gpurun_out/ncu_roles_$d.csv \
    python bench.py --steps 1 --warmup 1 --no-step-graph --no-cpu-baseline --no-e2e --no-extra --no-updates --no-secondary --no-small --no-parity-check > gpurun_out/ncu_roles_$d.log 2>&1
  echo "dbg $d rc $?"
done
