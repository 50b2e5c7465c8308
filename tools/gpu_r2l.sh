#!/bin/bash
TAG=${1:-r2l}
OUT=gpurun_out/$TAG
mkdir -p $OUT
for mb in 5 6; do for rf in 8 16 24 32; do
  if [ $mb = 6 ] && { [ $rf = 8 ] || [ $rf = 32 ]; }; then continue; fi
  OSPCUDA_PT_REFILL=$rf OSPCUDA_PT_TRACE_BLOCKS=$mb timeout 300 python bench.py --no-cpu-baseline --steps 20 > $OUT/bench_rf${rf}_mb$mb.json 2>> $OUT/bench.err
  python - <<PY
import json
d=json.load(open("$OUT/bench_rf${rf}_mb$mb.json"))
r=d["roofline"]
print("refill $rf blocks $mb: frame_ms", d["ms_per_step"], "trace", r["kernel_ms_per_frame"], r["other_classes_ms_per_frame"], "e2e ms", d["e2e"]["ms_per_step"])
PY
  OSPCUDA_PT_REFILL=$rf OSPCUDA_PT_TRACE_BLOCKS=$mb bash tools/gpu_wl.sh $TAG C4 C5s | grep -v rc=
done; done
