#!/bin/bash
# bench line + per-class ncu metrics of one C1 frame -> profiles/r2_c1_kernel_classes.json
TAG=${1:-r2o}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 600 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; tail -3 $OUT/bench.err
cat $OUT/bench.json
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum
timeout 900 ncu --metrics $M --clock-control none --cache-control none -c 400 --csv --log-file $OUT/c1_classes.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $OUT/c1_classes_under_ncu.log 2>&1; echo "ncu classes rc=$?"
python tools/ncu_classes.py $OUT/c1_classes.csv > $OUT/r2_c1_kernel_classes.json; cat $OUT/r2_c1_kernel_classes.json
