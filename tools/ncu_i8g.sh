#!/bin/bash
# Run on the GPU box: one `ncu --set full` capture of the dim-128 int8 GROUP pass and the group re-score at the c5 shard
# shape (12.5 M x 128, 1024 queries; tools/bench_shard.py).   usage: tools/ncu_i8g.sh TAG -> gpurun_out/TAG_c5_i8g.ncu-rep
set -u
TAG=${1:-i8g}
NCU="ncu --clock-control none --kernel-name-base demangled --set full --import-source on"
NCU=1 G=8 N=100000000 D=128 B=128 timeout -s KILL 400 $NCU -k 'regex:recall_scan_i8g_kernel|rescore_group_kernel' -s 10 -c 5 -f \
    -o gpurun_out/${TAG}_c5_i8g python tools/bench_shard.py > gpurun_out/${TAG}_ncu.log 2>&1
tail -3 gpurun_out/${TAG}_ncu.log
ls -la gpurun_out/${TAG}_c5_i8g*
