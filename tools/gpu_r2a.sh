#!/bin/bash
# Round-2 first pass: GPU parity suite + smoke + bench after the ray-queue wavefront / compact-hit traversal refactor.
TAG=${1:-r2a}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $OUT/gpu.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_gpu.log
tail -30 $OUT/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $OUT/smoke.log
tail -3 $OUT/smoke.log
timeout 600 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"
cat $OUT/bench.json; tail -5 $OUT/bench.err
timeout 900 python tools/bench_workloads.py C2 C3 C4 C5s --frames 3 --warmup 1 --out $OUT/workloads.json > $OUT/workloads.log 2>&1; echo "workloads rc=$?"
tail -8 $OUT/workloads.log
