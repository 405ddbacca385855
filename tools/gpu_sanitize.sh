#!/bin/bash
# Run on the GPU box: compute-sanitizer over the parity tests of the kernels written or changed in round 2 (small shapes only:
# the tools slow kernels down 10-100x).   usage: tools/gpu_sanitize.sh TAG -> gpurun_out/TAG_{memcheck,racecheck}.log
set -u
TAG=${1:-san}
SEL="tests/test_user_features_gpu.py::test_fm_with_user_fields_bit_exact tests/test_user_features_gpu.py::test_multi_head_score_map_and_rank_score tests/test_sort_dpp_gpu.py::test_dpp_hook_embeddings tests/test_sort_dpp_gpu.py::test_dpp_config4_shape tests/test_rank_gpu.py tests/test_mlp_gpu.py"
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 1 python -m pytest $SEL -m gpu -q --timeout 800 -x > gpurun_out/${TAG}_memcheck.log 2>&1
tail -5 gpurun_out/${TAG}_memcheck.log
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 1 python -m pytest "tests/test_group_gpu.py::test_group_equals_unsharded_path[2-64-24-300]" -m gpu -q --timeout 500 -x > gpurun_out/${TAG}_memcheck_group.log 2>&1
tail -4 gpurun_out/${TAG}_memcheck_group.log
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 1 python -m pytest tests/test_user_features_gpu.py::test_fm_with_user_fields_bit_exact tests/test_sort_dpp_gpu.py::test_dpp_hook_embeddings -m gpu -q --timeout 500 -x > gpurun_out/${TAG}_racecheck.log 2>&1
tail -5 gpurun_out/${TAG}_racecheck.log
