#!/bin/bash
# one sub-device's share of the C1 frame (tiles with tile % N == 0) on one GPU, tail start depth swept
TAG=${1:-r2s}
OUT=gpurun_out/$TAG
mkdir -p $OUT
for n in 8 4 2; do
  for tf in 2 3 4 5 6 7; do
    OSPRAY_NUM_SUBDEVICES=$n OSPRAY_SUBDEVICE_RANK=0 OSPCUDA_PT_TAIL_FROM=$tf timeout 300 python bench.py --no-cpu-baseline --steps 30 > $OUT/b.json 2>> $OUT/bench.err
    python - <<PY
import json
d=json.load(open("$OUT/b.json"))
c={e["class"]: e["ms_per_frame"] for e in d["roofline"]["classes"]}
print("share 1/$n tail from $tf: frame_ms", d["ms_per_step"], c)
PY
  done
done
