#!/bin/bash
# Round 2, call T (1 GPU): HCA encode variants (CRI_LIB_PATH): bench only.
set -u
OUT=gpurun_out
for v in "$@"; do
  LIBP=$PWD/pycricodecs_b200/libcricodecs_b200_$v.so
  [ $v = main ] && LIBP=$PWD/pycricodecs_b200/libcricodecs_b200.so
  CRI_LIB_PATH=$LIBP timeout 300 python bench.py --workload hca_encode --no-cpu --e2e-steps 0 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('hca_encode $v ms', round(d['ms_per_step'],3), d['parity_spot_check'])"
done
