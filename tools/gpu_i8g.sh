#!/bin/bash
# Run on the GPU box: parity tests of the int8 filter (both kernels), then the rank-0 recall part of an 8-way sharded c5
# step (12.5 M x 128, 1024 queries; tools/bench_shard.py) with the bf16 and the int8 index.  usage: tools/gpu_i8g.sh TAG
set -u
TAG=${1:-i8g}
timeout -s KILL 400 python -m pytest tests/test_recall_i8_gpu.py -m gpu -q -x --timeout 100 2>&1 | tail -30 > gpurun_out/${TAG}_pytest_i8.log
cat gpurun_out/${TAG}_pytest_i8.log
timeout -s KILL 400 python -m pytest tests/test_recall_gpu.py tests/test_shard_gpu.py tests/test_group_gpu.py -m gpu -q --timeout 150 2>&1 | tail -15 > gpurun_out/${TAG}_pytest_recall.log
cat gpurun_out/${TAG}_pytest_recall.log
for v in 0 1; do
  echo "== c5 shard shape (12.5 M x 128, 1024 queries), PRG_SCAN_INT8=$v"; PRG_SCAN_INT8=$v G=8 N=100000000 D=128 B=128 timeout -s KILL 300 python tools/bench_shard.py 2>&1 | tail -3
done > gpurun_out/${TAG}_shard_c5.log 2>&1
cat gpurun_out/${TAG}_shard_c5.log
