#!/bin/bash
# Round 2, call Q (1 GPU): ADX encode: tests, worker / mover time split (development build), bench.
set -u
OUT=gpurun_out
timeout 180 python -m pytest tests/test_adx_gpu.py tests/test_full_size_gpu.py -m gpu -x -q -k "adx" 2>&1 | tail -2
for v in "$@"; do
CRI_LIB_PATH=$PWD/pycricodecs_b200/libcricodecs_b200_$v.so timeout 120 python bench.py --workload adx_encode --no-cpu --steps 1 --warmup 1 --e2e-steps 0 > $OUT/r02q_$v.log 2>&1
echo "== $v"; grep "adx encode cta 0 group 0" $OUT/r02q_$v.log | sort | uniq | head -8
done
for w in adx_encode adx_decode; do
timeout 120 python bench.py --workload $w --no-cpu --e2e-steps 0 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$w ms', d['ms_per_step'], d['roofline']['frac'], d['parity_spot_check'])"
done
