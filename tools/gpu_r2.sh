#!/bin/bash
# Run on the GPU box (gpurun), round 2: bench lines of c4 (with the CPU baseline), c2, c3; the reference arm; launch list
# and one `ncu --set full` capture of the heavy kernels of a c4 step.   usage: tools/gpu_r2.sh TAG -> gpurun_out/TAG_*
set -u
TAG=${1:-r2}
timeout 600 python bench.py --steps 200 --warmup 5 > gpurun_out/${TAG}_bench_c4.log 2>&1
timeout 300 python bench.py --workload c2 --steps 200 --warmup 5 --no-cpu-baseline > gpurun_out/${TAG}_bench_c2.log 2>&1
timeout 300 python bench.py --workload c3 --steps 200 --warmup 5 --no-cpu-baseline > gpurun_out/${TAG}_bench_c3.log 2>&1
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/${TAG}_bench_ref.log 2>&1
NCU="ncu --clock-control none --kernel-name-base demangled"
timeout 400 $NCU -k regex:prg:: --metrics gpu__time_duration.sum -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-batcher > gpurun_out/${TAG}_launches_bench.log 2>&1
timeout 600 $NCU -k "regex:dpp_pair_kernel|gather_fm_kernel|mlp_layer_persistent_kernel|recall_scan_tc_kernel" --set full --import-source on \
    -s 24 -c 12 -f -o gpurun_out/${TAG}_full python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-batcher > gpurun_out/${TAG}_full_bench.log 2>&1
timeout 300 $NCU -k "regex:gather_fm_kernel" --set full -s 9 -c 2 -f -o gpurun_out/${TAG}_c3_full \
    python bench.py --workload c3 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_c3_full_bench.log 2>&1
python - <<PY
import json
for f in ("c4", "c2", "c3", "ref"):
    try:
        d = json.loads([x for x in open("gpurun_out/${TAG}_bench_%s.log" % f) if x.startswith("{")][-1])
        print(f, round(d["value"], 1), round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["value"], 1), d.get("stage_ms_per_step"),
              (d.get("roofline") or {}).get("frac"), (d.get("cpu_baseline") or {}).get("value"), d.get("dpp_bound", {}).get("frac"))
    except Exception as e:
        print(f, "no line:", e)
        print(open("gpurun_out/${TAG}_bench_%s.log" % f).read()[-1500:])
PY
