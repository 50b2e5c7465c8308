#!/bin/bash
TAG=${1:-r2r}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q -x > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_gpu.log
grep -E "^FAILED|passed|failed" $OUT/pytest_gpu.log | head
for cfg in "128 0" "128 1" "64 0" "86 0" "256 0"; do
  set -- $cfg
  OSPCUDA_PT_BATCH_TILES=$1 OSPCUDA_PT_ONE_LANE=$2 timeout 300 python bench.py --no-cpu-baseline --steps 30 > $OUT/b.json 2>> $OUT/bench.err
  python - <<PY
import json
d=json.load(open("$OUT/b.json"))
c={e["class"]: e["ms_per_frame"] for e in d["roofline"]["classes"]}
print("batch tiles $1 one-lane $2: frame_ms", d["ms_per_step"], c, "e2e", d["e2e"]["ms_per_step"])
PY
done
bash tools/gpu_wl.sh $TAG C4 C5s
OSPCUDA_PT_BATCH_TILES=128 bash tools/gpu_wl.sh $TAG C3
OSPCUDA_PT_BATCH_TILES=256 bash tools/gpu_wl.sh $TAG C3
