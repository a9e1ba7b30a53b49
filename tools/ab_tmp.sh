for n in "" _np _p7 _p8 _n7; do
  LM3D_LIB=$PWD/3d-localisation-and-mapping_b200/lm3d/liblm3d$n.so python bench.py --workload C3 --frames 1024 --steps 5 --warmup 3 --no-cpu --no-e2e --no-other --no-dropin 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); k=d['roofline']['kernel_ms']; print('lib$n', 'C3x1024', round(d['ms_per_step'],3), 'tile', round(k['tile_path'],3), 'blk', round(k['lift_block'],2))"
done
