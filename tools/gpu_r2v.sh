#!/bin/bash
# Round 2, call V (1 GPU): ADX tests + bench lines of both ADX workloads.
set -u
OUT=gpurun_out
timeout 300 python -m pytest tests/test_adx_gpu.py tests/test_full_size_gpu.py tests/test_wav_ingest.py -m gpu -x -q -k "adx or ingest or wav" 2>&1 | tail -2
for w in adx_encode adx_decode; do
timeout 120 python bench.py --workload $w --no-cpu --e2e-steps 0 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$w ms', d['ms_per_step'], d['roofline']['frac'], d['parity_spot_check'])"
done
