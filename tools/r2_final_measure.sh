#!/bin/bash
# dev helper (run under gpurun, one GPU): the round's N=1 evidence -- full bench line, launch lists, ncu --set full captures
set -x
O=gpurun_out
timeout 300 python -m pytest tests/test_gpu_tiles.py tests/test_gpu_fuzz.py tests/test_gpu_parity.py -m gpu -q 2>&1 | tail -3 > $O/r2_final_tests.log
(time timeout 900 python bench.py) > $O/r2_bench_n1.json 2> $O/r2_bench_n1.err
timeout 120 python bench.py --impl reference --steps 3 --warmup 1 > $O/r2_bench_reference_arm.json 2>> $O/r2_bench_n1.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"tile_|lift_|prep_|scale_boxes" -c 200 --csv --log-file $O/r2_launches_c2.csv python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e --no-other --no-dropin > $O/r2_ncu_c2.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"tile_|lift_|prep_|scale_boxes" -c 200 --csv --log-file $O/r2_launches_c3.csv python bench.py --workload C3 --frames 512 --steps 2 --warmup 3 --no-cpu --no-e2e > $O/r2_ncu_c3.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"tile_|lift_|prep_|scale_boxes" -c 200 --csv --log-file $O/r2_launches_c5.csv python bench.py --workload C5 --frames 256 --steps 2 --warmup 3 --no-cpu --no-e2e > $O/r2_ncu_c5.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:lift_quad_kernel -s 3 -c 1 -f -o $O/r2_prof_lift_quad python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e --no-other --no-dropin > $O/r2_prof_quad.log 2>&1
for k in tile_box tile_sum; do timeout 400 ncu --set full --clock-control none --import-source on -k regex:${k}_kernel -s 2 -c 1 -f -o $O/r2_prof_$k python bench.py --workload C3 --frames 64 --steps 1 --warmup 3 --no-e2e --no-cpu > $O/r2_prof_$k.log 2>&1; done
python tools/bench_cloud.py > $O/r2_frame_cloud.json 2>&1
python tools/bench_ingest.py > $O/r2_ingest.json 2>&1
python tools/bench_nms.py 4000 50 x > $O/r2_nms.json 2>&1
ls -la $O | tail -20
