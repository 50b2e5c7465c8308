#!/bin/bash
# A/B of whole library sets on C1 (alternating, same box): bash tools/ab_c1.sh build/lib_base ospray_b200/lib
for rep in 1 2; do for d in "$@"; do
  OSPRAY_B200_LIB_DIR=$PWD/$d timeout 300 python bench.py --no-cpu-baseline --steps 30 --warmup 5 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read()); c={x['class']:x['ms_per_frame'] for x in d['roofline']['classes']}
print('$d', d['ms_per_step'], c, 'e2e', d['e2e']['ms_per_step'])"
done; done
