"""Print a one-line digest of a bench.py JSON line (dev helper)."""
import json
import sys

for ln in open(sys.argv[1]):
    ln = ln.strip()
    if not ln.startswith("{"):
        continue
    d = json.loads(ln)
    r = d.get("roofline") or {}
    print(
        f"value={d['value']:.4g} {d['unit']} ms/step={d['ms_per_step']:.4g} frac={r.get('frac', 0):.4f} "
        f"kernel_ms={ {k: round(v, 4) for k, v in (r.get('kernel_ms') or {}).items()} } "
        f"e2e={(d.get('e2e') or {}).get('value')} match={(d.get('e2e') or {}).get('matches_device_path')} "
        f"cpu={(d.get('cpu_baseline') or {}).get('value')} rare={r.get('rare_paths')}"
    )
