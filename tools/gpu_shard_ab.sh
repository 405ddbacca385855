#!/bin/bash
# Run on the GPU box: rank-0 recall part of an 8-way sharded step at the c4 and c5 per-GPU shapes (tools/bench_shard.py),
# per-query (PRG_SCAN_GROUPS=0) against GROUP mode (=1) of the filter.   usage: tools/gpu_shard_ab.sh TAG
set -u
TAG=${1:-ab}
for g in 0 1; do
  echo "== c4 shard shape (1.25 M x 64, 512 queries), PRG_SCAN_GROUPS=$g"; PRG_SCAN_GROUPS=$g G=8 timeout 200 python tools/bench_shard.py 2>&1 | tail -3
done > gpurun_out/${TAG}_shard_c4.log 2>&1
cat gpurun_out/${TAG}_shard_c4.log
for g in 0 1; do
  echo "== c5 shard shape (12.5 M x 128, 1024 queries), PRG_SCAN_GROUPS=$g"; PRG_SCAN_GROUPS=$g G=8 N=100000000 D=128 B=128 timeout 300 python tools/bench_shard.py 2>&1 | tail -3
done > gpurun_out/${TAG}_shard_c5.log 2>&1
cat gpurun_out/${TAG}_shard_c5.log
