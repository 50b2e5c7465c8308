#!/bin/bash
TAG=${1:-r2h}; shift
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 900 python tools/bench_workloads.py ${@:-C2 C3 C4 C5s} --frames 3 --warmup 1 --out $OUT/workloads.json > $OUT/workloads.log 2>&1; echo "workloads rc=$?"
python - <<PY
import json
for l in open('$OUT/workloads.log'):
    try: d=json.loads(l)
    except Exception: continue
    if isinstance(d, dict): print(d.get('workload'), d.get('frame_ms_min'), d.get('mrays_per_s'), d.get('bvh_build_ms'), d.get('rays_split_primary_secondary_shadow'))
PY
