#!/bin/bash
# Run on the GPU box: per-kernel launch times (ncu --metrics gpu__time_duration.sum) of the rank-0 recall part of a sharded
# step at the c4 (G=8) and c5 shard shapes.   usage: tools/gpu_shard_launches.sh TAG
set -u
TAG=${1:-sl}
NCU="ncu --clock-control none --kernel-name-base demangled --metrics gpu__time_duration.sum --csv"
NCU=1 G=8 timeout -s KILL 200 $NCU --log-file gpurun_out/${TAG}_c4g8_launches.csv python tools/bench_shard.py > /dev/null 2>&1
NCU=1 G=8 N=100000000 D=128 B=128 timeout -s KILL 300 $NCU --log-file gpurun_out/${TAG}_c5g8_launches.csv python tools/bench_shard.py > /dev/null 2>&1
ls -la gpurun_out/${TAG}_*
