#!/bin/bash
# per-kernel time of the sample select at the single-GPU shape for the three CTA sizes (run on the GPU box)
for NT in 256 512 1024; do
  PRG_TOPR_NT=$NT NCU=1 G=1 timeout 200 ncu --clock-control none --kernel-name-base demangled -k regex:sample_topr \
    --metrics gpu__time_duration.sum --csv --log-file gpurun_out/topr_$NT.csv python tools/bench_shard.py > /dev/null 2>&1
  echo "NT=$NT: $(grep -o '"[0-9.]*"$' gpurun_out/topr_$NT.csv | tr -d '"' | tr '\n' ' ')"
done
