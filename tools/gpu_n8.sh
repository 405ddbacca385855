#!/bin/bash
# Run on an 8-GPU box (gpurun --gpus 8): the c5 step (100 M x 128 row-sharded, 1024 requests per step) and the c4 step over
# 8 ranks, one process per GPU over NCCL, as the driver launches them.   usage: tools/gpu_n8.sh TAG [N]
set -u
TAG=${1:-n8}
N=${2:-8}
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
if [ "$N" = "8" ]; then
timeout -s KILL 420 $RUN bench.py --gpus $N --workload c5 --steps 50 --warmup 5 > gpurun_out/${TAG}_c5.log 2>&1
grep '^{' gpurun_out/${TAG}_c5.log | tail -1 | cut -c1-1500
fi
timeout -s KILL 300 $RUN bench.py --gpus $N --steps 200 --warmup 5 > gpurun_out/${TAG}_c4.log 2>&1
grep '^{' gpurun_out/${TAG}_c4.log | tail -1 | cut -c1-1500
