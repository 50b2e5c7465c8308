#!/bin/bash
# warm-cache launch list of one workload: bash tools/gpu_ll.sh TAG WORKLOAD
TAG=$1; WL=$2
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 400 --csv --log-file $OUT/launches_$WL.csv \
    python tools/bench_workloads.py $WL --frames 2 --warmup 1 > $OUT/${WL}_under_ncu.log 2>&1; echo "ncu launches rc=$?"
python - <<PY
import csv, collections
lines=[l for l in open('$OUT/launches_$WL.csv') if not l.startswith('==')]
rows=[(x['Kernel Name'], float(x['Metric Value'])/1000.0) for x in csv.DictReader(lines)]
idx=[i for i,(k,_) in enumerate(rows) if 'ptGen' in k or 'renderFrameKernel' in k]
s=idx[-1]
tot=collections.Counter(); seq=[]
for k,v in rows[s:]:
    name=k.split('(')[0][-40:]
    tot[name]+=v; seq.append((name[-22:],round(v,1)))
print(seq[:50])
print(tot)
PY
