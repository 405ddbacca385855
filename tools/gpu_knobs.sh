#!/bin/bash
# Run on the GPU box: A/B bench runs over experiment knobs (environment variables read by the library).
# usage: tools/gpu_knobs.sh TAG "ENV1=a ENV2=b" "ENV1=c" ...   (an empty string = defaults)
set -u
TAG=$1; shift
i=0
for kv in "$@"; do
  env $kv timeout 200 python bench.py --steps 200 --warmup 5 --no-cpu-baseline --no-batcher > gpurun_out/${TAG}_$i.log 2>&1
  python - "$kv" gpurun_out/${TAG}_$i.log <<'PY'
import json, sys
try:
    d = json.loads([x for x in open(sys.argv[2]) if x.startswith("{")][-1])
    print(repr(sys.argv[1]), round(d["value"]), round(d["ms_per_step"], 4), round(d["e2e"]["value"]), {k: round(v, 4) for k, v in d["stage_ms_per_step"].items()})
except Exception as e:
    print(repr(sys.argv[1]), "no line:", e)
PY
  i=$((i+1))
done
