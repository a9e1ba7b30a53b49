#!/bin/bash
# dev helper (run under gpurun): ncu --set full of one lift_small launch -> gpurun_out/prof_small.ncu-rep
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:lift_small -s 3 -c 1 -f -o gpurun_out/prof_small \
    python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/ncu2.log 2>&1
echo "ncu rc=$?"
