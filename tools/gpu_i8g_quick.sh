#!/bin/bash
# quick check of the GROUP-mode recall paths: parity tests + the c5 / c4 shard lines.  usage: tools/gpu_i8g_quick.sh TAG
set -u
TAG=${1:-i8q}
timeout -s KILL 400 python -m pytest tests/test_recall_i8_gpu.py tests/test_recall_gpu.py tests/test_shard_gpu.py tests/test_group_gpu.py -m gpu -q -x --timeout 150 2>&1 | tail -5
for v in ${VARIANTS:-1}; do
  echo "== c5 shard shape (12.5 M x 128, 1024 queries), PRG_SCAN_INT8=$v"; PRG_SCAN_INT8=$v G=8 N=100000000 D=128 B=128 timeout -s KILL 300 python tools/bench_shard.py 2>&1 | tail -3
done > gpurun_out/${TAG}_shard_c5.log 2>&1
cat gpurun_out/${TAG}_shard_c5.log
echo "== c4 shard shape (1.25 M x 64, 512 queries)" > gpurun_out/${TAG}_shard_c4.log; G=8 timeout 200 python tools/bench_shard.py 2>&1 | tail -3 >> gpurun_out/${TAG}_shard_c4.log
cat gpurun_out/${TAG}_shard_c4.log
