#!/bin/bash
# dev helper (run here, after tools/r2_final_measure.sh came back): digest gpurun_out/r2_prof_<tag>.ncu-rep into
# profiles/r2_<tag>_ncu_summary.txt (key metrics) and profiles/r2_<tag>_per_line.txt (instruction / stall shares per source line)
# usage: tools/r2_digest.sh <tag> <kernel substring> "<command line the capture ran>"
TAG=$1; KERN=$2; CMD=$3
cd /root/repo/gpurun_out
REP=r2_prof_$TAG.ncu-rep
ncu -i $REP --page raw --csv 2>/dev/null > raw_$TAG.csv
ncu -i $REP --page source --csv --print-source sass 2>/dev/null > sass_$TAG.csv
ncu -i $REP --page source --csv --print-source cuda 2>/dev/null > src_$TAG.csv
SHA=$(cd /root/repo && python -c "import bench; print(bench.source_sha())")
OUT=/root/repo/profiles/r2_${TAG}_ncu_summary.txt
echo "# ncu --set full --clock-control none --import-source on -k regex:${KERN}, round-2 final build (csrc sha $SHA)" > $OUT
echo "# $CMD" >> $OUT
python - $TAG >> $OUT <<'PY'
import csv,sys
tag=sys.argv[1]
rows=list(csv.reader(open(f'raw_{tag}.csv')))
hdr=rows[0]; units=rows[1]; data=rows[2:]
keys=['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','launch__grid_size','launch__block_size','smsp__issue_active.avg.pct_of_peak_sustained_active','smsp__inst_executed.sum','smsp__thread_inst_executed_per_inst_executed.ratio','smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio','smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio','smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio','smsp__average_warps_issue_stalled_wait_per_issue_active.ratio','smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio','smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio','smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio','smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio','smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio','sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active','l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','l1tex__t_sector_hit_rate.pct','lts__t_sector_hit_rate.pct']
for k in keys:
    if k in hdr:
        i=hdr.index(k); print(k, units[i], data[0][i])
PY
mkdir -p /tmp/sass && cd /tmp/sass && rm -f *.cubin && cuobjdump -xelf all "/root/repo/3d-localisation-and-mapping_b200/lm3d/liblm3d.so" >/dev/null && nvdisasm --print-line-info lm3d_kernels.sm_100a.cubin > all.sass 2>/dev/null
cd /root/repo/gpurun_out
python /root/repo/tools/prof_lines.py $KERN sass_$TAG.csv src_$TAG.csv 45 > /root/repo/profiles/r2_${TAG}_per_line.txt 2>&1
head -12 $OUT; head -8 /root/repo/profiles/r2_${TAG}_per_line.txt
