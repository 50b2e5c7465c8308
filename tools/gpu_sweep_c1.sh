#!/bin/bash
# C1 sweep: tail start depth x SHADE resident CTAs (A/B on one box)
#   gpurun --timeout 600 -- 'bash tools/gpu_sweep_c1.sh r2z'
TAG=${1:-r2z}
OUT=gpurun_out/$TAG
mkdir -p $OUT
for tf in 5 6 7 8; do for sb in 4 5; do
  OSPCUDA_PT_TAIL_FROM=$tf OSPCUDA_PT_SHADE_BLOCKS=$sb timeout 120 python bench.py --no-cpu-baseline --steps 20 --warmup 3 > $OUT/b_${tf}_${sb}.json 2>/dev/null
  python - <<PY
import json
d=json.load(open("$OUT/b_${tf}_${sb}.json"))
c={x["class"]:x["ms_per_frame"] for x in d["roofline"]["classes"]}
print("tail_from $tf shade_blocks $sb frame_ms", d["ms_per_step"], c)
PY
done; done
