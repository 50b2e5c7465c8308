#!/bin/bash
TAG=${1:-r2q}
OUT=gpurun_out/$TAG
mkdir -p $OUT
for size in "512 256" "512 512" "1024 512" "1024 1024"; do
  set -- $size
    timeout 300 python bench.py --no-cpu-baseline --steps 30 --width $1 --height $2 > $OUT/b.json 2>> $OUT/bench.err
    python - <<PY
import json
d=json.load(open("$OUT/b.json"))
c={e["class"]: e["ms_per_frame"] for e in d["roofline"]["classes"]}
print("size $1x$2: frame_ms", d["ms_per_step"], c, "e2e", d["e2e"]["ms_per_step"])
PY
done
bash tools/gpu_wl.sh $TAG C4 C5s C3
