#!/bin/bash
# Second GPU pass: GPU parity suite, smoke, both bench arms, then the non-headline workloads (C2, C3, C5s)
# with one launch list + full capture of the C3 AO render kernel (the HBM/L2-bound traversal case).
#   gpurun --timeout 1700 -- 'bash tools/gpu_round2.sh r1b'
TAG=${1:-r1b}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $OUT/gpu.txt 2>&1
nproc > $OUT/host.txt; grep -m1 "model name" /proc/cpuinfo >> $OUT/host.txt
if [ "$2" != "notests" ]; then
  timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_gpu.log
  tail -5 $OUT/pytest_gpu.log
  timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $OUT/smoke.log
  tail -3 $OUT/smoke.log
fi
timeout 600 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"
cat $OUT/bench.json
timeout 900 python tools/bench_workloads.py C2 C3 C5s --frames 3 --warmup 1 --out $OUT/workloads.json > $OUT/workloads.log 2>&1; echo "workloads rc=$?"
cat $OUT/workloads.log | tail -5
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $OUT/launches_c3.csv \
    python tools/bench_workloads.py C3 --frames 1 --warmup 1 > $OUT/c3_under_ncu.log 2>&1; echo "ncu c3 launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:renderFrameKernel -s 1 -c 1 \
    -o $OUT/c3_ao_full -f python tools/bench_workloads.py C3 --frames 1 --warmup 1 > $OUT/ncu_c3_full.log 2>&1; echo "ncu c3 full rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:renderFrameKernel -s 1 -c 1 \
    -o $OUT/c2_scivis_full -f python tools/bench_workloads.py C2 --frames 1 --warmup 1 > $OUT/ncu_c2_full.log 2>&1; echo "ncu c2 full rc=$?"
ls -la $OUT
