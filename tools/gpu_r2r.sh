#!/bin/bash
# Round 2, call R (1 GPU): HCA encode with counted bit costs + single-stream MDCT butterflies: parity, bench, ncu.
set -u
TAG=${1:-r02r}
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests/test_hca_encode_gpu.py tests/test_wav_ingest.py tests/test_full_size_gpu.py tests/test_regressions_gpu.py tests/test_usm_audio.py -m gpu -x -q > $OUT/${TAG}_pytest_encode.log 2>&1
tail -4 $OUT/${TAG}_pytest_encode.log
timeout 300 python bench.py --workload hca_encode --no-cpu --e2e-steps 1 > $OUT/${TAG}_bench_hca_encode.json 2> $OUT/${TAG}_bench_hca_encode.err
tail -2 $OUT/${TAG}_bench_hca_encode.err
python -c "
import json; d = json.load(open('$OUT/${TAG}_bench_hca_encode.json')); print('hca_encode ms', round(d['ms_per_step'], 3), d['parity_spot_check'], d['config'].get('streams_per_gpu'))"
if [ "${2:-}" = ncu ]; then
timeout 600 ncu --set full --clock-control none --import-source on -k regex:hca_encode_kernel -s 4 -c 1 -o $OUT/${TAG}_prof_hca_encode -f \
    python bench.py --workload hca_encode --steps 1 --warmup 3 --no-cpu --e2e-steps 1 > $OUT/${TAG}_ncu_encode.log 2>&1
fi
ls -la $OUT | grep ${TAG}
