#!/bin/bash
# Round 2, call AQ (1 GPU): device-pointer calls after the WAV fetch change (tests), then the encoders' e2e_device.
set -u
timeout 600 python -m pytest tests/test_device_api_gpu.py tests/test_wav_ingest.py tests/test_sharding_gpu.py -m gpu -x -q 2>&1 | tail -2
for w in adx_encode hca_encode; do
  timeout 300 python bench.py --workload $w --no-cpu --no-companion --steps 3 --warmup 3 --e2e-steps 3 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$w dev ms', round(d['e2e_device']['ms_per_step'],2), d['e2e_device']['matches_host_path'], 'kernels', round(d['ms_per_step'],2))"
done
