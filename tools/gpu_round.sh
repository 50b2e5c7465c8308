#!/bin/bash
# One GPU-box pass: GPU parity suite, smoke, bench (both arms), ncu launch list + one full capture of the
# dominant kernel. Everything lands in gpurun_out/<tag>/ (tag = $1, default "r1").
#   gpurun --timeout 1500 -- 'bash tools/gpu_round.sh r1 [notests]'
TAG=${1:-r1}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $OUT/gpu.txt 2>&1
if [ "$2" != "notests" ]; then
  timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_gpu.log
  tail -5 $OUT/pytest_gpu.log
  timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $OUT/smoke.log
  tail -3 $OUT/smoke.log
fi
timeout 600 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"
cat $OUT/bench.json
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > $OUT/bench_ref.json 2> $OUT/bench_ref.err; echo "ref rc=$?"
cat $OUT/bench_ref.json
# launch list (cold-cache, serialised: shares only)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/bench_under_ncu.log 2>&1; echo "ncu launches rc=$?"
# full capture of the dominant kernel class (first-bounce TRACE launch of a warm frame)
KERNEL=${NCU_KERNEL:-ptTraceKernel}
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$KERNEL -s ${NCU_SKIP:-40} -c ${NCU_COUNT:-2} \
    -o $OUT/${KERNEL}_full -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/ncu_full.log 2>&1; echo "ncu full rc=$?"
ls -la $OUT
