#!/bin/bash
# proxy for the per-rank frame of C1 at N = 8 / 4 / 2: 1/8, 1/4, 1/2 of the pixels on one GPU, tail start depth swept
TAG=${1:-r2p}
OUT=gpurun_out/$TAG
mkdir -p $OUT
for size in "512 256" "512 512" "1024 512"; do
  set -- $size
  for tf in 1 2 3 4 5 7; do
    OSPCUDA_PT_TAIL_FROM=$tf timeout 300 python bench.py --no-cpu-baseline --steps 30 --width $1 --height $2 > $OUT/b.json 2>> $OUT/bench.err
    python - <<PY
import json
d=json.load(open("$OUT/b.json"))
c={e["class"]: e["ms_per_frame"] for e in d["roofline"]["classes"]}
print("size $1x$2 tail from $tf: frame_ms", d["ms_per_step"], c)
PY
  done
done
