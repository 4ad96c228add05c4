#!/bin/bash
# Round 2, call AF (1 GPU): pipelined device-pointer batch calls: tests that use them, phases (CRI_TRACE=1), e2e_device per workload.
set -u
OUT=gpurun_out
timeout 900 python -m pytest tests/test_device_api_gpu.py tests/test_sharding_gpu.py tests/test_hca_crypt_gpu.py tests/test_regressions_gpu.py tests/test_full_size_gpu.py -m gpu -x -q 2>&1 | tail -3
for w in hca_decode adx_encode; do
  echo "== $w"
  CRI_TRACE=1 timeout 300 python bench.py --workload $w --no-cpu --no-companion --steps 3 --warmup 3 --e2e-steps 2 2>&1 >/dev/null | grep "cri trace" | tail -1 | cut -c1-600
done
for p in 1 2 4 8; do
  CRI_DEV_PIECES=$p timeout 300 python bench.py --no-cpu --no-companion --steps 3 --warmup 3 --e2e-steps 3 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('hca_decode pieces $p dev ms', round(d['e2e_device']['ms_per_step'],2), d['e2e_device']['matches_host_path'])"
done
for w in adx_encode adx_decode hca_encode hca_decrypt hca_decrypt_decode; do
  timeout 300 python bench.py --workload $w --no-cpu --no-companion --steps 3 --warmup 3 --e2e-steps 3 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$w dev ms', round(d['e2e_device']['ms_per_step'],2), d['e2e_device']['matches_host_path'], 'kernels', round(d['ms_per_step'],2))"
done
