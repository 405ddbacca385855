#!/bin/bash
# Run on the GPU box: compute-sanitizer memcheck over parity tests of the int8 / GROUP-mode filter kernels and the group
# re-score (recall_i8.cu, recall.cu).   usage: tools/gpu_sanitize_i8.sh TAG -> gpurun_out/TAG_memcheck_i8.log
set -u
TAG=${1:-san}
T=tests/test_recall_i8_gpu.py
SEL="$T::test_int8_filter_matches_oracle[300001-5-200] $T::test_dim128_group_passes_use_the_int8_index[262399-65-50] $T::test_dim64_group_pass_kernels_agree[129] $T::test_nan_inf_zero_rows_and_negative_thresholds $T::test_dim128_negative_thresholds_nan_rows_and_padding_groups"
timeout -s KILL 280 compute-sanitizer --tool memcheck --error-exitcode 1 python -m pytest $SEL -m gpu -q --timeout 250 -x > gpurun_out/${TAG}_memcheck_i8.log 2>&1
tail -6 gpurun_out/${TAG}_memcheck_i8.log
