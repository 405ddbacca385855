#!/bin/bash
# Run on the GPU box: one `ncu --set full` capture of the int8 filter pass at the C2 shape (10 M x 64, 64 queries).
# usage: tools/ncu_i8.sh TAG -> gpurun_out/TAG_scan_i8.ncu-rep
set -u
TAG=${1:-i8}
NCU="ncu --clock-control none --kernel-name-base demangled --set full --import-source on"
timeout -s KILL 250 $NCU -k 'regex:recall_scan_i8_kernel' -s 4 -c 1 -f -o gpurun_out/${TAG}_scan_i8 python tools/bench_recall.py > gpurun_out/${TAG}_ncu.log 2>&1
tail -3 gpurun_out/${TAG}_ncu.log
timeout 200 python tools/bench_recall.py 2>&1 | tail -3
ls -la gpurun_out/${TAG}_scan_i8*
