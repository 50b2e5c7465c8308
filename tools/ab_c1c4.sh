#!/bin/bash
# A/B of whole library sets on C1 (bench.py) and C4 / C5s (workload bench): bash tools/ab_c1c4.sh build/lib_A ospray_b200/lib
for rep in 1 2; do for d in "$@"; do
  OSPRAY_B200_LIB_DIR=$PWD/$d timeout 300 python bench.py --no-cpu-baseline --steps 30 --warmup 5 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read()); c={x['class']:x['ms_per_frame'] for x in d['roofline']['classes']}
print('$d C1', d['ms_per_step'], c)"
  OSPRAY_B200_LIB_DIR=$PWD/$d python tools/bench_workloads.py C4 C5s --frames 3 --warmup 1 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('$d', d['workload'], d['frame_ms_min'])"
done; done
