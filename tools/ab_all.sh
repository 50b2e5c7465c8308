# usage: bash tools/ab_all.sh dir1 dir2 ...  : cornell bench line + C3/C5s workloads per library directory
for d in "$@"; do echo "LIB=$d"; OSPRAY_B200_LIB_DIR=$PWD/$d timeout 300 python bench.py --no-cpu-baseline --steps 20 --warmup 5 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('C1', d['ms_per_step'], d['value'], d['roofline']['kernel_ms_per_frame'], d['roofline']['other_classes_ms_per_frame'])"
OSPRAY_B200_LIB_DIR=$PWD/$d timeout 600 python tools/bench_workloads.py ${WL:-C3 C5s} --frames 3 --warmup 1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(d['workload'], 'frame_ms', d['frame_ms'], 'min', d['frame_ms_min'], 'Mrays/s', d['mrays_per_s'])"; done
