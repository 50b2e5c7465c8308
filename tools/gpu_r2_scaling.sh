#!/bin/bash
# round-2 scaling on one 8-GPU box: C1 and C5 at N = 1, 2, 4, 8 through bench.py exactly as the driver launches it
#   gpurun --gpus 8 --timeout 1500 -- 'bash tools/gpu_r2_scaling.sh'
OUT=gpurun_out/r2scale
mkdir -p $OUT
port=29600
for wl in C1 C5; do
  steps=30; [ $wl = C5 ] && steps=6
  for n in 1 2 4 8; do
    port=$((port+1))
    if [ $n = 1 ]; then
      python bench.py --gpus 1 --workload $wl --steps $steps --warmup 5 --no-cpu-baseline > $OUT/${wl}_n$n.json 2> $OUT/${wl}_n$n.err
    else
      python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port bench.py --gpus $n --workload $wl --steps $steps --warmup 5 --no-cpu-baseline > $OUT/${wl}_n$n.json 2> $OUT/${wl}_n$n.err
    fi
    echo "$wl N=$n rc=$? $(tail -1 $OUT/${wl}_n$n.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['e2e']['ms_per_step'])" 2>&1)"
  done
done
