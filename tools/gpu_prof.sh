#!/bin/bash
# ncu --set full of named kernels in the default bench; usage: bash tools/gpu_prof.sh <tag> <workload> <regex> [<regex> ...]
TAG=$1; WL=$2; shift 2
OUT=gpurun_out; mkdir -p $OUT
for K in "$@"; do
  N=$(echo $K | tr -c 'a-zA-Z0-9_' '_')
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$K -s 4 -c 1 -o $OUT/${TAG}_prof_$N -f \
      python bench.py --workload $WL --steps 1 --warmup 3 --no-cpu --e2e-steps 1 > $OUT/${TAG}_ncu_$N.log 2>&1
  tail -2 $OUT/${TAG}_ncu_$N.log
done
