#!/bin/bash
# Run on the GPU box (gpurun): launch list of the bench command + one `ncu --set full` capture of every kernel of a step.
# Outputs land in gpurun_out/ (copy the summaries you keep into profiles/).
set -u
TAG=${1:-step}
NCU="ncu --clock-control none --kernel-name-base demangled -k regex:prg::"
timeout 600 $NCU --metrics gpu__time_duration.sum -s 47 -c 90 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-batcher > gpurun_out/${TAG}_launches_bench.log 2>&1
timeout 900 $NCU --set full --import-source on -s 47 -c 15 -f -o gpurun_out/${TAG}_full \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-batcher > gpurun_out/${TAG}_full_bench.log 2>&1
ls -la gpurun_out/
