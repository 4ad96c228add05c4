#!/bin/bash
# Round 2, call P (1 GPU): ncu capture of one kernel (argument: kernel regex, workload, tag)
set -u
K=$1; W=$2; TAG=$3
OUT=gpurun_out
timeout 600 ncu --set full --import-source on --clock-control none -k regex:$K -c 1 -s 2 -o $OUT/${TAG} python bench.py --workload $W --no-cpu --steps 1 --warmup 1 --e2e-steps 0 > $OUT/${TAG}.log 2>&1
tail -2 $OUT/${TAG}.log
