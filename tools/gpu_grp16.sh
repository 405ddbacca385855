#!/bin/bash
# Run on the GPU box: GROUP-mode parity tests, then the rank-0 recall part of a sharded c4 step at G = 2, 4, 8 with
# recall_tc.cu's group pass (PRG_SCAN_GRP16=0) and the 16-epilogue-warp one (=1).   usage: tools/gpu_grp16.sh TAG
set -u
TAG=${1:-g16}
timeout -s KILL 500 python -m pytest tests/test_recall_i8_gpu.py tests/test_recall_gpu.py tests/test_shard_gpu.py tests/test_group_gpu.py -m gpu -q -x --timeout 150 2>&1 | tail -5
for G in 8 4 2; do for v in 0 1; do
  echo "== c4 shard shape G=$G, PRG_SCAN_GRP16=$v"; PRG_SCAN_GRP16=$v G=$G timeout -s KILL 200 python tools/bench_shard.py 2>&1 | tail -3
done; done > gpurun_out/${TAG}_shard_c4.log 2>&1
cat gpurun_out/${TAG}_shard_c4.log
