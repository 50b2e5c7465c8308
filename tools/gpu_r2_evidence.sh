#!/bin/bash
# evidence pass: per-class ncu metrics of one C3 frame (wavefront ao: DRAM bytes against the algorithmic bytes, lanes, issue)
# and of one C1 frame; compute-sanitizer memcheck over a slice of the parity suite
#   gpurun --timeout 1500 -- 'bash tools/gpu_r2_evidence.sh r2ev'
TAG=${1:-r2ev}
OUT=gpurun_out/$TAG
mkdir -p $OUT
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum
timeout 900 ncu --metrics $M --clock-control none --cache-control none -c 300 --csv --log-file $OUT/c3_classes.csv \
    python tools/bench_workloads.py C3 --frames 2 --warmup 0 > $OUT/c3_under_ncu.log 2>&1; echo "ncu c3 rc=$?"
python tools/ncu_classes.py $OUT/c3_classes.csv aoGenKernel > $OUT/r2_c3_kernel_classes.json; head -c 1500 $OUT/r2_c3_kernel_classes.json
timeout 600 ncu --metrics $M --clock-control none --cache-control none -c 400 --csv --log-file $OUT/c1_classes.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $OUT/c1_under_ncu.log 2>&1; echo "ncu c1 rc=$?"
python tools/ncu_classes.py $OUT/c1_classes.csv > $OUT/r2_c1_kernel_classes.json
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_parity.py tests/test_gpu_textures.py tests/test_edge_cases.py -m gpu -q -x \
    -k "cornell_pt or textured_pt or emissive_pt or gravity_pt or spheres_ao or streamlines_pt or debug_Ng or RGBA8 or SRGB or cutout or instanced_scivis or volume_opaque" > $OUT/sanitizer_memcheck.log 2>&1; echo "sanitizer rc=$?"; tail -5 $OUT/sanitizer_memcheck.log
