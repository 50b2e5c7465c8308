#!/bin/bash
# full capture of one kernel launch of a workload: bash tools/gpu_ncu_wl.sh TAG WORKLOAD KERNEL_REGEX SKIP
TAG=$1; WL=$2; KR=$3; SKIP=${4:-1}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$KR -s $SKIP -c 1 \
    -o $OUT/${WL}_${KR}_full -f python tools/bench_workloads.py $WL --frames 1 --warmup 0 > $OUT/ncu_${WL}_${KR}.log 2>&1; echo "ncu rc=$?"
