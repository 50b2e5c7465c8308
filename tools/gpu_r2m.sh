#!/bin/bash
TAG=${1:-r2m}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_gpu.log
grep -E "^FAILED|passed|failed" $OUT/pytest_gpu.log | head -20
bash tools/gpu_wl.sh $TAG C3
OSPCUDA_AO_MEGAKERNEL=1 bash tools/gpu_wl.sh ${TAG}_mega C3
