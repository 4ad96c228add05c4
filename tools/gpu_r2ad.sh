#!/bin/bash
# Round 2, call AD (1 GPU): planners after the host-side changes: tests that plan, then the CRI_TRACE phases again.
set -u
timeout 600 python -m pytest tests/test_adx_gpu.py tests/test_hca_crypt_gpu.py tests/test_device_api_gpu.py tests/test_wav_ingest.py tests/test_regressions_gpu.py tests/test_dropin_gpu.py -m gpu -x -q 2>&1 | tail -2
for w in adx_encode hca_decrypt hca_decode; do
  echo "== $w"
  CRI_TRACE=1 timeout 300 python bench.py --workload $w --no-cpu --no-companion --steps 3 --warmup 3 --e2e-steps 2 2>&1 >/dev/null | grep "cri trace" | tail -2
done
