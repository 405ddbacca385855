#!/bin/bash
# Run on the GPU box: one `ncu --set full` capture of the filter pass (256 queries per pass, threshold mode) at the
# c4 / c5 per-GPU shard shapes of an 8-way sharded run, emulated on one GPU by tools/bench_shard.py.
# usage: tools/ncu_shard_scan.sh TAG  -> gpurun_out/TAG_scan_c4g8.ncu-rep, gpurun_out/TAG_scan_c5g8.ncu-rep
set -u
TAG=${1:-scan}
NCU="ncu --clock-control none --kernel-name-base demangled --set full --import-source on"
NCU=1 G=8 timeout 250 $NCU -k 'regex:recall_scan_tc_kernel<\(int\)[0-9]+, \(int\)4' -s 18 -c 2 -f -o gpurun_out/${TAG}_scan_c4g8 \
    python tools/bench_shard.py > gpurun_out/${TAG}_c4.log 2>&1
tail -2 gpurun_out/${TAG}_c4.log
NCU=1 G=8 N=100000000 D=128 B=128 timeout 300 $NCU -k 'regex:recall_scan_tc_kernel<\(int\)[0-9]+, \(int\)4' -s 36 -c 2 -f -o gpurun_out/${TAG}_scan_c5g8 \
    python tools/bench_shard.py > gpurun_out/${TAG}_c5.log 2>&1
tail -2 gpurun_out/${TAG}_c5.log
ls -la gpurun_out/${TAG}_scan*
