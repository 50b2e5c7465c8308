#!/bin/bash
# full captures of the depth-1 TRACE and depth-1 SHADE launches of a C1 frame (+ launch list)
TAG=${1:-r2d}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $OUT/launches_c1.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $OUT/c1_under_ncu.log 2>&1; echo "ncu launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ptTraceKernel -s 1 -c 1 \
    -o $OUT/pt_trace_full -f python bench.py --steps 1 --warmup 0 --no-cpu-baseline > $OUT/ncu_trace.log 2>&1; echo "ncu trace rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ptShadeKernel -s 1 -c 1 \
    -o $OUT/pt_shade_full -f python bench.py --steps 1 --warmup 0 --no-cpu-baseline > $OUT/ncu_shade.log 2>&1; echo "ncu shade rc=$?"
