#!/bin/bash
# Run on the GPU box: parity tests of the int8 filter index, then the c4 step with and without it.  usage: tools/gpu_i8.sh TAG
set -u
TAG=${1:-i8}
timeout -s KILL 300 python -m pytest tests/test_recall_i8_gpu.py -m gpu -q -x --timeout 100 2>&1 | tail -30 > gpurun_out/${TAG}_pytest_i8.log
cat gpurun_out/${TAG}_pytest_i8.log
timeout -s KILL 300 python -m pytest tests/test_recall_gpu.py tests/test_snapshot_gpu.py -m gpu -q --timeout 100 2>&1 | tail -15 > gpurun_out/${TAG}_pytest_recall.log
cat gpurun_out/${TAG}_pytest_recall.log
for v in 0 1; do
  PRG_SCAN_INT8=$v timeout -s KILL 200 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --no-batcher > gpurun_out/${TAG}_bench_int8_$v.log 2>&1
  python - <<PY
import json
try:
    d = json.loads([x for x in open("gpurun_out/${TAG}_bench_int8_$v.log") if x.startswith("{")][-1])
    print("int8=$v", round(d["value"]), round(d["ms_per_step"], 4), d["stage_ms_per_step"], "e2e", round(d["e2e"]["value"]), d["roofline"]["frac"], d["roofline"]["launch_ms"])
except Exception as e:
    print("int8=$v no line:", e); print(open("gpurun_out/${TAG}_bench_int8_$v.log").read()[-1500:])
PY
done
