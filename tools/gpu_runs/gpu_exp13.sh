#!/bin/bash
# development: per-kernel durations of the split stage B (few-slice launch)
set -x
mkdir -p gpurun_out
P="python bench.py --workload C5 --batch 8 --steps 1 --warmup 1 --no-cpu --no-e2e"
$P > gpurun_out/plain13.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches13.csv $P > gpurun_out/ncu13.log 2>&1
python - <<'PY'
import csv,collections
rows=list(csv.reader(open("gpurun_out/launches13.csv")))
hi=next(i for i,r in enumerate(rows) if r and r[0]=="ID")
ix={h:i for i,h in enumerate(rows[hi])}
agg=collections.OrderedDict()
for r in rows[hi+1:]:
    if len(r)<len(rows[hi]): continue
    n=r[ix["Kernel Name"]].split("(")[0][:40]
    a=agg.setdefault(n,[0,0.0]); a[0]+=1; a[1]+=float(r[ix["Metric Value"]])
for n,a in agg.items(): print(n,a[0],round(a[1]/1e6,2),"ms")
PY
