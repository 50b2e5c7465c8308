#!/bin/bash
# end-of-round evidence on one B200: the whole GPU suite, smoke, the bench line (own arm and reference arm), the launch list of
# the bench command, the other workloads
#   gpurun --timeout 2400 -- 'bash tools/gpu_r2_final.sh r2final'
TAG=${1:-r2final}
OUT=gpurun_out/$TAG
mkdir -p $OUT
python -m pytest tests -m gpu -q -x > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $OUT/smoke.log 2>&1; tail -1 $OUT/smoke.log
python bench.py > $OUT/r2_bench_c1_n1.json 2> $OUT/bench.err; echo "bench rc=$?"; head -c 600 $OUT/r2_bench_c1_n1.json; echo
python bench.py --impl reference --steps 3 --warmup 1 > $OUT/r2_bench_c1_reference_arm.json 2> $OUT/bench_ref.err; echo "ref arm rc=$?"; head -c 400 $OUT/r2_bench_c1_reference_arm.json; echo
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/r2_launches_bench_c1.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/bench_under_ncu.log 2>&1; echo "ncu launch list rc=$?"
python tools/bench_workloads.py C2 C3 C4 C5s > $OUT/r2_workloads.json 2> $OUT/workloads.err; echo "workloads rc=$?"; cat $OUT/r2_workloads.json | head -c 1500
python tools/bench_materials.py 2>/dev/null > $OUT/r2_material_frames.json; echo "materials rc=$?"; cut -c1-70 $OUT/r2_material_frames.json
