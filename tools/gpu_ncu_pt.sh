#!/bin/bash
# ncu full captures of the wavefront PT kernel classes on the C1 bench workload (warm frame): TRACE(d=1), SHADE(d=1)
TAG=${1:-r1f}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 600 python bench.py --no-cpu-baseline > $OUT/bench.json 2> $OUT/bench.err; cat $OUT/bench.json
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ptTraceKernel -s 40 -c 2 -o $OUT/ptTrace_full -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/ncu_trace.log 2>&1; echo "ncu trace rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ptShadeKernel -s 36 -c 2 -o $OUT/ptShade_full -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/ncu_shade.log 2>&1; echo "ncu shade rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/bench_under_ncu.log 2>&1; echo "ncu launches rc=$?"
ls -la $OUT
