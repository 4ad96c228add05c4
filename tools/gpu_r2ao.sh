#!/bin/bash
# Round 2, call AO (1 GPU): A/B of an ADX variant against the committed build in one session (tests on the variant first).
set -u
V=${1:-mm3}
CRI_LIB_PATH=$PWD/pycricodecs_b200/libcricodecs_b200_$V.so timeout 600 python -m pytest tests/test_adx_gpu.py tests/test_full_size_gpu.py -m gpu -x -q -k "adx" 2>&1 | tail -2
for v in main $V main $V; do
  LIBP=$PWD/pycricodecs_b200/libcricodecs_b200_$v.so; [ $v = main ] && LIBP=$PWD/pycricodecs_b200/libcricodecs_b200.so
  for w in adx_decode adx_encode; do
  CRI_LIB_PATH=$LIBP timeout 300 python bench.py --workload $w --no-cpu --e2e-steps 0 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$w $v ms', round(d['ms_per_step'],4), d['parity_spot_check'])"
  done
done
