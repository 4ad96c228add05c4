#!/bin/bash
# Round 2, call AH (1 GPU): tests over the device-pointer calls, then the piece count sweep for HCA decode.
set -u
timeout 900 python -m pytest tests/test_device_api_gpu.py tests/test_sharding_gpu.py tests/test_regressions_gpu.py tests/test_hca_decode_gpu.py tests/test_awb.py tests/test_acb.py -m gpu -x -q 2>&1 | tail -2
for p in 2 4 6 8 12; do
  CRI_DEV_PIECES=$p timeout 300 python bench.py --no-cpu --no-companion --steps 3 --warmup 3 --e2e-steps 4 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('hca_decode pieces $p dev ms', round(d['e2e_device']['ms_per_step'],2), d['e2e_device']['matches_host_path'])"
done
