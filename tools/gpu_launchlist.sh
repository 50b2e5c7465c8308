#!/bin/bash
# launch list of the C1 bench with warm caches (--cache-control none): per-launch durations close to the un-profiled run
TAG=${1:-r2f}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 300 --csv --log-file $OUT/launches_c1_warm.csv \
    python bench.py --steps 3 --warmup 2 --no-cpu-baseline > $OUT/c1_under_ncu.log 2>&1; echo "ncu launches rc=$?"
