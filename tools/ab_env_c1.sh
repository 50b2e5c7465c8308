#!/bin/bash
# A/B of an environment knob on C1 with one library set (alternating): bash tools/ab_env_c1.sh OSPCUDA_PT_FOLD_IN_TRACE 0 1
VAR=$1; shift
for rep in 1 2; do for v in "$@"; do
  env $VAR=$v timeout 300 python bench.py --no-cpu-baseline --steps 30 --warmup 5 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read()); c={x['class']:x['ms_per_frame'] for x in d['roofline']['classes']}
print('$VAR=$v', d['ms_per_step'], c, 'e2e', d['e2e']['ms_per_step'], 'rays', d['config']['rays_per_frame'])"
done; done
