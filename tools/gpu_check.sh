#!/bin/bash
# Run on the GPU box (gpurun): GPU parity tests, smoke, the bench line (chained launches on / off) and the launch list.
# usage: tools/gpu_check.sh TAG   -> gpurun_out/TAG_*.log
set -u
TAG=${1:-check}
timeout 420 python -m pytest tests -m gpu -q --timeout 180 2>&1 | tail -15 > gpurun_out/${TAG}_pytest.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1
timeout 300 python bench.py --steps 200 --warmup 5 > gpurun_out/${TAG}_bench.log 2>&1
[ -n "${SKIP_NOPDL:-}" ] || PRG_PDL=0 timeout 200 python bench.py --steps 200 --warmup 5 --no-cpu-baseline --no-batcher > gpurun_out/${TAG}_bench_nopdl.log 2>&1
NCU="ncu --clock-control none --kernel-name-base demangled -k regex:prg::"
timeout 300 $NCU --metrics gpu__time_duration.sum -s 40 -c 100 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-batcher > gpurun_out/${TAG}_launches_bench.log 2>&1
tail -3 gpurun_out/${TAG}_pytest.log
python - <<PY
import json
for f in ("gpurun_out/${TAG}_bench.log", "gpurun_out/${TAG}_bench_nopdl.log"):
    try:
        l = [x for x in open(f) if x.startswith("{")][-1]
        d = json.loads(l)
        print(f, d["value"], d["ms_per_step"], d["e2e"]["value"], d["gpu_launches"], d["stage_ms_per_step"])
    except Exception as e:
        print(f, "no line:", e)
PY
