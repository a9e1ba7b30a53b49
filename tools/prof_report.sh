#!/bin/bash
# dev helper (run here): digest gpurun_out/prof_small.ncu-rep: key metrics + per-source-line instruction counts
REP=${1:-/root/repo/gpurun_out/prof_small.ncu-rep}; SLOTS=${2:-21612497}; KERN=${3:-lift_small}
cd /root/repo/gpurun_out
ncu -i $REP --page raw --csv 2>/dev/null > raw.csv
ncu -i $REP --page source --csv --print-source sass 2>/dev/null > sass.csv
python - <<'PY'
import csv
rows=list(csv.reader(open('/root/repo/gpurun_out/raw.csv')))
hdr=rows[0]; units=rows[1]; data=rows[2:]
keys=['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','smsp__issue_active.avg.pct_of_peak_sustained_active','smsp__inst_executed.sum','smsp__thread_inst_executed_per_inst_executed.ratio','launch__grid_size','smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio','smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio','smsp__average_warps_issue_stalled_wait_per_issue_active.ratio','smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio','smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio','smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio','smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio','smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio','sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active','l1tex__t_sector_hit_rate.pct','lts__t_sector_hit_rate.pct']
for k in keys:
    if k in hdr:
        i=hdr.index(k); print(k, units[i], [r[i] for r in data])
PY
mkdir -p /tmp/sass && cd /tmp/sass && rm -f *.cubin && cuobjdump -xelf all "/root/repo/3d-localisation-and-mapping_b200/lm3d/liblm3d.so" >/dev/null && nvdisasm --print-line-info lm3d_kernels.sm_100a.cubin > all.sass 2>/dev/null
python /root/repo/tools/prof_join.py $KERN /root/repo/gpurun_out/sass.csv $SLOTS 2>&1 | awk '$3+0 >= 0.7 || /total|other/' | cut -c1-200
