# usage: bash tools/ab_workloads.sh "C2 C3" dir1 dir2 ...   (device-timed workload numbers per library directory)
WL=$1; shift
for d in "$@"; do echo "LIB=$d"; OSPRAY_B200_LIB_DIR=$PWD/$d timeout 600 python tools/bench_workloads.py $WL --frames 4 --warmup 2 ${BOXES:+--boxes $BOXES} | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(d['workload'], 'frame_ms', d['frame_ms'], 'min', d['frame_ms_min'], 'Mrays/s', d['mrays_per_s'], 'build_ms', d['bvh_build_ms'])"; done
