#!/bin/bash
# the driver's round-end sequence: GPU tests, smoke, both bench arms (+ workloads)
TAG=${1:-r2t}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_gpu.log
grep -E "^FAILED|passed|failed" $OUT/pytest_gpu.log | head
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $OUT/smoke.log
timeout 600 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; cat $OUT/bench.json | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('C1 frame_ms', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], {e['class']: e['ms_per_frame'] for e in d['roofline']['classes']}, d['roofline'].get('fb_ops_4k'))"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_ref.json 2>> $OUT/bench.err; echo "ref rc=$?"; cat $OUT/bench_ref.json | cut -c1-300
bash tools/gpu_wl.sh $TAG C2 C3 C4 C5s
