#!/bin/bash
# Run on the GPU box: one `ncu --set full` capture of the recall kernels at the C2 shape (10 M x 64, 64 queries): sample
# scan, threshold, int8 filter pass, exact re-score, refine select.   usage: tools/ncu_i8.sh TAG -> gpurun_out/TAG_recall_i8.ncu-rep
set -u
TAG=${1:-i8}
NCU="ncu --clock-control none --kernel-name-base demangled --set full --import-source on"
timeout -s KILL 280 $NCU -k 'regex:recall_scan_i8_kernel|rescore_kernel|refine_select_kernel|tilemax_tau_kernel|recall_scan_tc_kernel' -s 20 -c 5 -f \
    -o gpurun_out/${TAG}_recall_i8 python tools/bench_recall.py > gpurun_out/${TAG}_ncu.log 2>&1
tail -3 gpurun_out/${TAG}_ncu.log
ls -la gpurun_out/${TAG}_recall_i8*
