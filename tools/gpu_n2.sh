#!/bin/bash
# two GPUs: the in-process multidevice over NVLink peer memory + the torchrun bench arm
TAG=${1:-r2u}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi -L | tee $OUT/gpus.txt
timeout 900 python -m pytest tests/test_gpu_multidevice.py tests/test_gpu_peer_exchange.py tests/test_gpu_fb_ops.py -m gpu -q > $OUT/pytest_multi.log 2>&1; echo "pytest rc=$?"; tail -4 $OUT/pytest_multi.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 30 --warmup 5 > $OUT/bench_n2.json 2> $OUT/bench_n2.err; echo "bench n2 rc=$?"
python -c "
import json
d=json.load(open('$OUT/bench_n2.json'))
print('N=2 frame_ms', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], d['config']['partition'][:60])"
timeout 600 python tools/bench_multidevice.py 1 2 > $OUT/multidevice.log 2>&1; echo "multidevice bench rc=$?"; cat $OUT/multidevice.log | tail -8
