#!/bin/bash
# per-launch device times of the lock-step kernels for one sweep of a config
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:dense_ --csv --log-file gpurun_out/dense_launches.csv python scripts/prof_sweep.py ${1:-c2} 1 >/dev/null 2>&1
python - <<EOF
import csv, collections
rows=[r for r in csv.reader(open("gpurun_out/dense_launches.csv")) if len(r)>10 and r[0].isdigit()]
agg=collections.OrderedDict()
for r in rows:
    k=r[4].split("(")[0]; agg.setdefault(k,[]).append(float(r[-1])/1e3)
for k,v in agg.items(): print(f"{k:28s} n={len(v):3d} total {sum(v):9.1f} us  each "+" ".join(f"{x:.0f}" for x in v[:12]))
EOF
