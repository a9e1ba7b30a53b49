#!/bin/bash
# dev helper (run under gpurun): a few scheduler / cache metrics of one lift_small launch for each library variant given
M=gpu__time_duration.sum,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,sm__icc_request_hit_rate.pct,sm__icc_requests.sum,gcc__cache_requests_type_instruction.sum.pct_of_peak_sustained_elapsed,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio,smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio,smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio,smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio,smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio
for v in "$@"; do
  [ "$v" = "base" ] && v=""
  export LM3D_LIB=$PWD/3d-localisation-and-mapping_b200/lm3d/liblm3d$v.so
  timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/plain.log 2>&1 && \
  ncu --metrics $M --clock-control none -k regex:lift_small -s 3 -c 1 --csv --log-file gpurun_out/m$v.csv python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/ncu3.log 2>&1
  echo "variant [$v] rc=$?"
done
