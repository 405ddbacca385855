#!/bin/bash
# Run on the GPU box (gpurun): the opt-in experiments that were written but not yet run on a B200.
#   1. csrc/dpp_pair.cu  — DPP in one wave (2-CTA clusters, features partly in tensor memory): parity test, then bench
#   2. PRG_GATHER_HINTS=1 — L2 cache hints on the gather's loads: bench
#   3. PRG_FAST_SORT=1    — one-element-per-thread bitonic sort (csrc/bitonic.cuh) in the score sort and the recall's
#                           refine select: the sort / recall / fused-path parity tests under the knob, then bench
#   5. config scan128_nqb  — 128 / 256 queries per filter pass at dim 128 (C5 shape): gated parity test (run with 1.)
#   4. PRG_RECALL_TILEMAX=1 — recall threshold from per-tile maxima of the sample (no 40 MB of sample keys, no top-r
#                           select): recall / fused-path / full-size parity tests under the knob, then bench
# usage: tools/gpu_experiments.sh TAG   -> gpurun_out/TAG_*
set -u
TAG=${1:-exp}
PRG_TEST_EXPERIMENTAL=1 timeout 300 python -m pytest tests/test_sort_dpp_gpu.py tests/test_recall_gpu.py -m gpu -q --timeout 200 -k "pair or dim128" 2>&1 | tail -30 > gpurun_out/${TAG}_pytest_pair.log
tail -5 gpurun_out/${TAG}_pytest_pair.log
PRG_FAST_SORT=1 timeout 400 python -m pytest tests/test_sort_dpp_gpu.py tests/test_recall_gpu.py tests/test_pipeline_gpu.py -m gpu -q --timeout 200 2>&1 | tail -12 > gpurun_out/${TAG}_pytest_fastsort.log
tail -3 gpurun_out/${TAG}_pytest_fastsort.log
PRG_RECALL_TILEMAX=1 timeout 500 python -m pytest tests/test_recall_gpu.py tests/test_pipeline_gpu.py tests/test_full_size_gpu.py tests/test_shard_gpu.py -m gpu -q --timeout 300 2>&1 | tail -12 > gpurun_out/${TAG}_pytest_tilemax.log
tail -3 gpurun_out/${TAG}_pytest_tilemax.log
tools/gpu_knobs.sh ${TAG}_knob "PRG_RECALL_TILEMAX=1" "PRG_RECALL_TILEMAX=1 PRG_DPP_PAIR=1 PRG_GATHER_HINTS=1 PRG_FAST_SORT=1"
tools/gpu_knobs.sh ${TAG}_knob2 "" "PRG_DPP_PAIR=1" "PRG_GATHER_HINTS=1" "PRG_FAST_SORT=1" "PRG_DPP_PAIR=1 PRG_GATHER_HINTS=1 PRG_FAST_SORT=1"
