import csv,sys
for v in sys.argv[1:]:
    v = "" if v=="base" else v
    rows=[r for r in csv.reader(open(f'/root/repo/gpurun_out/m{v}.csv')) if len(r)>10]
    print("==",v or "base", "  ".join(f"{r[-3].replace('smsp__average_warps_issue_stalled_','st_').replace('_per_issue_active.ratio','').replace('.avg.pct_of_peak_sustained_active','').replace('.sum.pct_of_peak_sustained_elapsed','')}={r[-1]}" for r in rows[1:]))
