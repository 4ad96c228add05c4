#!/bin/bash
# Round 2, call N (1 GPU): ADX kernels: tests, where the decode worker's time goes (clock64 instrumentation build), timings.
set -u
TAG=${1:-r02n}
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python -m pytest tests/test_adx_gpu.py tests/test_full_size_gpu.py -m gpu -x -q -k "adx" > $OUT/${TAG}_pytest_adx.log 2>&1
tail -3 $OUT/${TAG}_pytest_adx.log
CRI_LIB_PATH=$PWD/pycricodecs_b200/libcricodecs_b200_adxtime.so timeout 300 python bench.py --workload adx_decode --no-cpu --steps 1 --warmup 1 --e2e-steps 0 > $OUT/${TAG}_adx_timing.log 2>&1
grep "adx decode cta" $OUT/${TAG}_adx_timing.log | sort | uniq -c | sort -rn | head -12
for w in adx_decode adx_encode; do
timeout 300 python bench.py --workload $w --no-cpu --e2e-steps 0 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$w ms', d['ms_per_step'], d['roofline']['frac'], d['parity_spot_check'])"
done
