#!/bin/bash
TAG=${1:-r2n}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_gpu.log
grep -E "^FAILED|passed|failed" $OUT/pytest_gpu.log | head -20
timeout 300 python bench.py --no-cpu-baseline --steps 20 > $OUT/bench.json 2>> $OUT/bench.err
python - <<PY
import json
d=json.load(open("$OUT/bench.json"))
r=d["roofline"]
print("C1 frame_ms", d["ms_per_step"], "trace", r["kernel_ms_per_frame"], r["other_classes_ms_per_frame"], "e2e ms", d["e2e"]["ms_per_step"])
PY
bash tools/gpu_wl.sh $TAG C2 C3 C4 C5s
