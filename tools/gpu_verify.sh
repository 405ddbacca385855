#!/bin/bash
# Run on the GPU box (gpurun): verification of HEAD without profiling — GPU parity tests, smoke, the default bench line.
# usage: tools/gpu_verify.sh TAG   -> gpurun_out/TAG_*
set -u
TAG=${1:-verify}
timeout -s KILL 420 python -m pytest tests -m gpu -q --timeout 180 2>&1 | tail -25 > gpurun_out/${TAG}_pytest.log
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1
timeout 240 python bench.py --steps 200 --warmup 5 > gpurun_out/${TAG}_bench.log 2>&1
tail -4 gpurun_out/${TAG}_pytest.log; tail -2 gpurun_out/${TAG}_smoke.log; tail -1 gpurun_out/${TAG}_bench.log | cut -c1-400
