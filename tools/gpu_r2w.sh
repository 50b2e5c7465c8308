#!/bin/bash
# C2 after the iterator fix: volume parity tests, A/B old / new / 128-register variant at 128^3 and 512^3, full ncu at 512^3
TAG=${1:-r2w}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q -k "volume or gravity or scivis" > $OUT/pytest_vol.log 2>&1; echo "pytest rc=$?"; tail -2 $OUT/pytest_vol.log
for lib in build/lib_old ospray_b200/lib build/lib_mb8; do
  for dim in 128 512; do
    n=$(basename $lib)
    OSPRAY_B200_LIB_DIR=$PWD/$lib timeout 600 python tools/bench_workloads.py C2 --dim $dim --frames 4 --warmup 1 --out $OUT/c2_${n}_$dim.json > $OUT/c2_${n}_$dim.log 2>&1
    python - <<PY
import json
d=json.load(open("$OUT/c2_${n}_$dim.json"))[0]
print("$lib", $dim, d["frame_ms_all"], d["mean_rgb"])
PY
  done
done
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:renderFrameKernel -s 1 -c 1 \
    -o $OUT/c2_512_scivis_full -f python tools/bench_workloads.py C2 --dim 512 --frames 1 --warmup 1 > $OUT/ncu_c2_512.log 2>&1; echo "ncu rc=$?"
