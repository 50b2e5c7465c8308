#!/bin/bash
# Final single-GPU pass of round 1 (after the treelet BVH, the tie rule, the peer exchange and the debug renderer):
# suite, smoke, both bench arms, launch list, ncu captures of the wavefront TRACE kernel (C1) and the AO megakernel
# (C3, treelet BVH), device-timed C2/C3/C4/C5s.   gpurun --timeout 1500 -- 'bash tools/gpu_round_r1final.sh r1f'
TAG=${1:-r1f}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $OUT/gpu.txt 2>&1
nproc > $OUT/host.txt; grep -m1 "model name" /proc/cpuinfo >> $OUT/host.txt
timeout 600 python -m pytest tests -m gpu -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_gpu.log; tail -2 $OUT/pytest_gpu.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $OUT/smoke.log; tail -2 $OUT/smoke.log
timeout 400 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; cut -c1-300 $OUT/bench.json
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_ref.json 2> $OUT/bench_ref.err; echo "ref rc=$?"; cut -c1-200 $OUT/bench_ref.json
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/launches_c1.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/bench_under_ncu.log 2>&1; echo "ncu launches rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ptTraceKernel -s 40 -c 2 -o $OUT/ptTrace_full -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/ncu_trace.log 2>&1; echo "ncu trace rc=$?"
timeout 600 python tools/bench_workloads.py C2 C3 C4 C5s --frames 4 --warmup 2 --out $OUT/workloads.json > $OUT/workloads.log 2>&1; echo "workloads rc=$?"; grep '^{' $OUT/workloads.log | cut -c1-330
timeout 600 ncu --set full --clock-control none --import-source on -k regex:renderFrameKernel -s 1 -c 1 -o $OUT/c3_ao_full -f python tools/bench_workloads.py C3 --frames 1 --warmup 1 > $OUT/ncu_c3_full.log 2>&1; echo "ncu c3 full rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $OUT/launches_c3.csv python tools/bench_workloads.py C3 --frames 1 --warmup 1 > $OUT/c3_under_ncu.log 2>&1; echo "ncu c3 launches rc=$?"
BOXES= bash tools/ab_treelet.sh "C3 C5s" 0 2 3 > $OUT/ab_treelet.txt 2>&1; grep -v "ospcuda bvh" $OUT/ab_treelet.txt
ls -la $OUT
