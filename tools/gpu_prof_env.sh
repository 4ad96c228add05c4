#!/bin/bash
# usage: bash tools/gpu_prof_env.sh <tag> <kernel regex> "ENV=.." ["ENV=.." ...]  -> one ncu --set full capture per environment
TAG=$1; K=$2; shift 2
OUT=gpurun_out; mkdir -p $OUT
i=0
for E in "$@"; do
  i=$((i+1))
  env $E timeout 900 ncu --set full --clock-control none --import-source on -k regex:$K -s 4 -c 1 -o $OUT/${TAG}_prof_$i -f \
      python bench.py --steps 1 --warmup 3 --no-cpu --e2e-steps 1 > $OUT/${TAG}_ncu_$i.log 2>&1
  tail -1 $OUT/${TAG}_ncu_$i.log
done
