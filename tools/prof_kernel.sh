#!/bin/bash
# dev helper (run under gpurun): ncu --set full of one launch of kernel $1 (regex) -> gpurun_out/prof_$2.ncu-rep
K=${1:-lift_tma}; TAG=${2:-tma}
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e > gpurun_out/plain_$TAG.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:$K -s 3 -c 1 -f -o gpurun_out/prof_$TAG \
    python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_$TAG.log 2>&1
echo "ncu rc=$?"
