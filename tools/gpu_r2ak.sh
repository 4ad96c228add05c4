#!/bin/bash
# Round 2, call AK (1 GPU): the device-API tests with the new pieces test; then everything once more from the clean build.
set -u
timeout 600 python -m pytest tests/test_device_api_gpu.py -m gpu -x -q 2>&1 | tail -3
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('default bench', round(d['value']/1e6,1), 'M frames/s', round(d['ms_per_step'],3), 'frac', round(d['roofline']['frac'],4), 'e2e', round(d['e2e']['ms_per_step'],1), 'dev', round(d['e2e_device']['ms_per_step'],2), 'launches', d['gpu_launches'], d['clocks']['reasons'], 'enc', round(d['hca_encode']['ms_per_step'],2), 'adx', round(d['adx_encode']['ms_per_step'],2))"
