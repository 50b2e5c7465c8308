#!/bin/bash
# C2 at 512^3 (the HBM-sized volume): frame time + full ncu capture of the scivis marcher
TAG=${1:-r2v}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 900 python tools/bench_workloads.py C2 --dim 512 --frames 3 --warmup 1 --out $OUT/c2_512.json > $OUT/c2_512.log 2>&1; echo "c2 512 rc=$?"; tail -2 $OUT/c2_512.log
timeout 900 python tools/bench_workloads.py C2 --dim 256 --frames 3 --warmup 1 --out $OUT/c2_256.json > $OUT/c2_256.log 2>&1; echo "c2 256 rc=$?"; tail -1 $OUT/c2_256.log
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:renderFrameKernel -s 1 -c 1 \
    -o $OUT/c2_512_scivis_full -f python tools/bench_workloads.py C2 --dim 512 --frames 1 --warmup 1 > $OUT/ncu_c2_512.log 2>&1; echo "ncu rc=$?"
