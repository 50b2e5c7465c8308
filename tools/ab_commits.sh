# A/B of whole library sets built from other commits (git worktree under build/wt_<commit>, `make` inside):
#   gpurun -- 'bash tools/ab_commits.sh "C3" build/wt_f6e4122/ospray_b200/lib ospray_b200/lib'
WL=$1; shift
for d in "$@"; do echo "LIB=$d"; OSPRAY_B200_LIB_DIR=$PWD/$d timeout 300 python tools/bench_workloads.py $WL --frames 3 --warmup 1 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(d['workload'], 'frame_ms', d['frame_ms'], 'all', d.get('frame_ms_all'), 'min', d['frame_ms_min'], 'Mrays/s', d['mrays_per_s'], 'rays', d['rays_per_frame'])
    elif 'Error' in l or 'error' in l: print(l.strip()[:200])"; done
