#!/bin/bash
TAG=${1:-r2e}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_gpu.log
tail -12 $OUT/pytest_gpu.log
for sb in 4 5 6; do
  OSPCUDA_PT_SHADE_BLOCKS=$sb timeout 300 python bench.py --no-cpu-baseline --steps 20 > $OUT/bench_sb${sb}.json 2>> $OUT/bench.err
  python - <<PY
import json
d=json.load(open("$OUT/bench_sb${sb}.json"))
r=d["roofline"]
print("shade blocks $sb: frame_ms", d["ms_per_step"], "trace", r["kernel_ms_per_frame"], r["other_classes_ms_per_frame"], "e2e ms", d["e2e"]["ms_per_step"])
PY
done
