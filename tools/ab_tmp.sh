for w in "C3 2000" "C5 1000"; do set -- $w
  python bench.py --workload $1 --frames $2 --steps 5 --warmup 3 --no-cpu --no-e2e --no-other --no-dropin 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); k=d['roofline']['kernel_ms']; print('$1', round(d['ms_per_step'],2), 'frac', round(d['roofline']['frac'],4), 'tile', round(k['tile_path'],2), 'blk', round(k['lift_block'],2), 'handed', d['roofline']['rare_paths']['tile_path_handed_to_block'])"
done
