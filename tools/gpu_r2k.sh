#!/bin/bash
TAG=${1:-r2k}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_gpu.log
tail -5 $OUT/pytest_gpu.log
for rf in 4 8 16; do
  OSPCUDA_PT_REFILL=$rf timeout 300 python bench.py --no-cpu-baseline --steps 20 > $OUT/bench_rf${rf}.json 2>> $OUT/bench.err
  python - <<PY
import json
d=json.load(open("$OUT/bench_rf${rf}.json"))
r=d["roofline"]
print("refill $rf: frame_ms", d["ms_per_step"], "trace", r["kernel_ms_per_frame"], r["other_classes_ms_per_frame"], "e2e ms", d["e2e"]["ms_per_step"])
PY
done
bash tools/gpu_wl.sh $TAG C3 C4 C5s
