#!/bin/bash
# Full GPU pass for the round record: parity suite, smoke, both bench arms, launch list + ncu captures of the wavefront
# PT kernels (C1), device-timed C2/C3/C5s, launch list + ncu capture of the C3 AO megakernel and the C2 volume kernel.
#   gpurun --timeout 2400 -- 'bash tools/gpu_round_final.sh r1z'
TAG=${1:-r1z}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $OUT/gpu.txt 2>&1
nproc > $OUT/host.txt; grep -m1 "model name" /proc/cpuinfo >> $OUT/host.txt
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_gpu.log; tail -3 $OUT/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $OUT/smoke.log; tail -2 $OUT/smoke.log
timeout 600 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; cat $OUT/bench.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_ref.json 2> $OUT/bench_ref.err; echo "ref rc=$?"; cat $OUT/bench_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/launches_c1.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/bench_under_ncu.log 2>&1; echo "ncu launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ptTraceKernel -s 40 -c 2 -o $OUT/ptTrace_full -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/ncu_trace.log 2>&1; echo "ncu trace rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ptShadeKernel -s 36 -c 2 -o $OUT/ptShade_full -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/ncu_shade.log 2>&1; echo "ncu shade rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ptFusedKernel -s 6 -c 1 -o $OUT/ptFused_full -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/ncu_fused.log 2>&1; echo "ncu fused rc=$?"
timeout 900 python tools/bench_workloads.py C2 C3 C5s --frames 4 --warmup 2 --out $OUT/workloads.json > $OUT/workloads.log 2>&1; echo "workloads rc=$?"; tail -3 $OUT/workloads.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $OUT/launches_c3.csv python tools/bench_workloads.py C3 --frames 1 --warmup 1 > $OUT/c3_under_ncu.log 2>&1; echo "ncu c3 launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:renderFrameKernel -s 1 -c 1 -o $OUT/c3_ao_full -f python tools/bench_workloads.py C3 --frames 1 --warmup 1 > $OUT/ncu_c3_full.log 2>&1; echo "ncu c3 full rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:renderFrameKernel -s 1 -c 1 -o $OUT/c2_scivis_full -f python tools/bench_workloads.py C2 --frames 1 --warmup 1 > $OUT/ncu_c2_full.log 2>&1; echo "ncu c2 full rc=$?"
ls -la $OUT
