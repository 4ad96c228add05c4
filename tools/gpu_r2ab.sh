#!/bin/bash
# Round 2, call AB (1 GPU): encoder after the phase-boundary fix: parity, racecheck + memcheck over the encoder tests, bench.
set -u
TAG=${1:-r02ab}; OUT=gpurun_out; mkdir -p $OUT
timeout 600 python -m pytest tests/test_hca_encode_gpu.py tests/test_wav_ingest.py tests/test_full_size_gpu.py tests/test_usm_audio.py -m gpu -x -q 2>&1 | tail -2
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 1 python -m pytest tests/test_hca_encode_gpu.py tests/test_adx_gpu.py -m gpu -x -q -k "not looping" > $OUT/${TAG}_sanitizer_racecheck.log 2>&1
echo "racecheck rc=$?" >> $OUT/${TAG}_sanitizer_racecheck.log; tail -3 $OUT/${TAG}_sanitizer_racecheck.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 1 python -m pytest tests/test_hca_encode_gpu.py tests/test_wav_ingest.py -m gpu -x -q > $OUT/${TAG}_sanitizer_memcheck.log 2>&1
echo "memcheck rc=$?" >> $OUT/${TAG}_sanitizer_memcheck.log; tail -3 $OUT/${TAG}_sanitizer_memcheck.log
timeout 300 python bench.py --workload hca_encode --no-cpu --e2e-steps 0 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('hca_encode ms', round(d['ms_per_step'],3), d['parity_spot_check'])"
