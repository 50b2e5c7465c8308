#!/bin/bash
# C3 / C4 against the TRACE knobs (resident CTAs per SM, refill threshold) on one box
for cfg in "5 16" "6 16" "6 24" "5 24" "5 8"; do set -- $cfg
  OSPCUDA_PT_TRACE_BLOCKS=$1 OSPCUDA_PT_REFILL=$2 python tools/bench_workloads.py C3 C4 --frames 3 --warmup 1 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('blocks $1 refill $2', d['workload'], d['frame_ms_min'])"
done
