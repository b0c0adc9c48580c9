#!/bin/bash
# per-kernel device time of one batched topN call (c5s shape)
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/topn_launches.csv python bench.py --config c5s --steps 1 --warmup 0 >/dev/null 2>&1
python - <<EOF
import csv, collections
rows=[r for r in csv.reader(open("gpurun_out/topn_launches.csv")) if len(r)>10 and r[0].isdigit()]
agg=collections.OrderedDict()
for r in rows:
    k=r[4].split("(")[0][:60]; agg.setdefault(k,[]).append(float(r[-1])/1e3)
tot=sum(sum(v) for v in agg.values())
for k,v in sorted(agg.items(), key=lambda kv:-sum(kv[1])): print(f"{k:62s} n={len(v):4d} total {sum(v)/1e3:9.2f} ms  ({100*sum(v)/tot:4.1f}%)")
print("sum of kernels", tot/1e3, "ms")
EOF
