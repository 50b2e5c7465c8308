#!/bin/bash
TAG=${1:-r2g}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_gpu.log
tail -5 $OUT/pytest_gpu.log
for tf in 0 5 6 7 8 9 10; do
  OSPCUDA_PT_TAIL_FROM=$tf timeout 300 python bench.py --no-cpu-baseline --steps 20 > $OUT/bench_tf${tf}.json 2>> $OUT/bench.err
  python - <<PY
import json
d=json.load(open("$OUT/bench_tf${tf}.json"))
r=d["roofline"]
print("tail from $tf: frame_ms", d["ms_per_step"], "trace", r["kernel_ms_per_frame"], r["other_classes_ms_per_frame"], "e2e ms", d["e2e"]["ms_per_step"])
PY
done
