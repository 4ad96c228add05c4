#!/bin/bash
# Round 2, call I (1 GPU): ADX encode with the first pass moved to the mover warps.
set -u
TAG=${1:-r02i}
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests/test_adx_gpu.py tests/test_wav_ingest.py tests/test_full_size_gpu.py tests/test_device_api_gpu.py tests/test_regressions_gpu.py tests/test_batch_pipeline_gpu.py tests/test_dropin_gpu.py -m gpu -x -q > $OUT/${TAG}_pytest_adx.log 2>&1
tail -6 $OUT/${TAG}_pytest_adx.log
timeout 300 python bench.py --workload adx_encode --no-cpu --e2e-steps 2 > $OUT/${TAG}_bench_adx_encode.json 2> $OUT/${TAG}_bench_adx_encode.err
tail -2 $OUT/${TAG}_bench_adx_encode.err
python -c "
import json; d = json.load(open('$OUT/${TAG}_bench_adx_encode.json')); print('adx_encode ms', round(d['ms_per_step'], 3), 'frac', round(d['roofline']['frac'], 4), d['parity_spot_check'])"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:adx_encode_fast -s 4 -c 1 -o $OUT/${TAG}_prof_adx_encode -f \
    python bench.py --workload adx_encode --steps 1 --warmup 3 --no-cpu --e2e-steps 1 > $OUT/${TAG}_ncu_adx_encode.log 2>&1
ls -la $OUT | grep ${TAG}
