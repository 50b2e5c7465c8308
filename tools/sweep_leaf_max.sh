for m in ${LEAVES:-1 2 3}; do echo "LEAF_MAX=$m"; OSPCUDA_LEAF_MAX=$m timeout 600 python tools/bench_workloads.py C3 C5s --frames 3 --warmup 1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(d['workload'], 'frame_ms', d['frame_ms'], 'min', d['frame_ms_min'], 'Mrays/s', d['mrays_per_s'], 'build_ms', d['bvh_build_ms'])"; done
