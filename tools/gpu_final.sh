#!/bin/bash
# Run on the GPU box (gpurun): the round's final verification — GPU parity tests (batcher tests also with the pipelined
# turn taking off), smoke, the bench lines (c4 default, c2, c3), the launch list and one `ncu --set full` capture of a step.
# usage: tools/gpu_final.sh TAG   -> gpurun_out/TAG_*
set -u
TAG=${1:-final}
timeout -s KILL 600 python -m pytest tests -m gpu -q --timeout 180 2>&1 | tail -25 > gpurun_out/${TAG}_pytest.log
PRG_BATCHER_PIPELINE=0 timeout 200 python -m pytest tests/test_batcher_gpu.py -m gpu -q --timeout 120 2>&1 | tail -8 > gpurun_out/${TAG}_pytest_batcher_nopipe.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1
timeout 300 python bench.py --steps 200 --warmup 5 > gpurun_out/${TAG}_bench.log 2>&1
timeout 200 python bench.py --workload c2 --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/${TAG}_bench_c2.log 2>&1
timeout 200 python bench.py --workload c3 --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/${TAG}_bench_c3.log 2>&1
NCU="ncu --clock-control none --kernel-name-base demangled -k regex:prg::"
timeout 300 $NCU --metrics gpu__time_duration.sum -s 40 -c 100 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-batcher > gpurun_out/${TAG}_launches_bench.log 2>&1
timeout 400 $NCU --set full --import-source on -s 44 -c 14 -f -o gpurun_out/${TAG}_full \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-batcher > gpurun_out/${TAG}_full_bench.log 2>&1
tail -4 gpurun_out/${TAG}_pytest.log; tail -2 gpurun_out/${TAG}_pytest_batcher_nopipe.log; tail -2 gpurun_out/${TAG}_smoke.log
python - <<PY
import json
for f in ("gpurun_out/${TAG}_bench.log", "gpurun_out/${TAG}_bench_c2.log", "gpurun_out/${TAG}_bench_c3.log"):
    try:
        d = json.loads([x for x in open(f) if x.startswith("{")][-1])
        b = d.get("e2e_batcher") or {}
        print(f, round(d["value"]), round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["value"]), "batcher", round(b.get("value", 0)),
              b.get("p50_ms"), b.get("p99_ms"), d["roofline"]["kernel"], round(d["roofline"]["frac"], 3), d["gpu_launches"])
    except Exception as e:
        print(f, "no line:", e)
PY
