# A/B of the treelet (SAH refinement) passes of the BVH builder: device-timed workloads per OSPCUDA_TREELET_PASSES value.
#   gpurun -- 'bash tools/ab_treelet.sh "C1 C3 C5s C4" 0 2 3'
WL=$1; shift
for p in "$@"; do echo "TREELET_PASSES=$p"; OSPCUDA_BVH_STATS=1 OSPCUDA_TREELET_PASSES=$p timeout 900 python tools/bench_workloads.py $WL --frames 4 --warmup 2 ${BOXES:+--boxes $BOXES} 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('[ospcuda bvh]'): print('   ', l.strip())
    if l.startswith('{'):
        d=json.loads(l); print(d['workload'], 'frame_ms', d['frame_ms'], 'min', d['frame_ms_min'], 'Mrays/s', d['mrays_per_s'], 'build_ms', d['bvh_build_ms'], 'mean_rgb', d['mean_rgb'])"; done
