#!/bin/bash
# Round 2, call AG (1 GPU): device-pointer batch calls after the piece policy / host-thread granularity change.
set -u
timeout 900 python -m pytest tests/test_device_api_gpu.py tests/test_sharding_gpu.py tests/test_regressions_gpu.py tests/test_cabi_gpu.py -m gpu -x -q 2>&1 | tail -2
CRI_TRACE=1 timeout 300 python bench.py --no-cpu --no-companion --steps 3 --warmup 3 --e2e-steps 2 2>&1 >/dev/null | grep "cri trace" | tail -1 | cut -c1-900
for w in hca_decode adx_encode adx_decode hca_encode hca_decrypt hca_decrypt_decode; do
  timeout 300 python bench.py --workload $w --no-cpu --no-companion --steps 3 --warmup 3 --e2e-steps 3 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$w dev ms', round(d['e2e_device']['ms_per_step'],2), d['e2e_device']['matches_host_path'], 'kernels', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['ms_per_step'],1), round(d['e2e']['frac_of_pcie_ceiling'],3))"
done
