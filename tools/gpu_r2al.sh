#!/bin/bash
# Round 2, call AL (1 GPU): sanitizer over the final build: racecheck (encoders, ADX, crypt), memcheck (device-pointer calls in pieces, encoders, decode).
set -u
TAG=${1:-r02al}; OUT=gpurun_out; mkdir -p $OUT
timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 1 python -m pytest tests/test_hca_encode_gpu.py tests/test_adx_gpu.py tests/test_hca_crypt_gpu.py -m gpu -x -q -k "not looping" > $OUT/${TAG}_sanitizer_racecheck.log 2>&1
echo "racecheck rc=$?" >> $OUT/${TAG}_sanitizer_racecheck.log; tail -3 $OUT/${TAG}_sanitizer_racecheck.log
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 1 python -m pytest tests/test_device_api_gpu.py tests/test_hca_encode_gpu.py tests/test_hca_decode_gpu.py tests/test_wav_ingest.py -m gpu -x -q > $OUT/${TAG}_sanitizer_memcheck.log 2>&1
echo "memcheck rc=$?" >> $OUT/${TAG}_sanitizer_memcheck.log; tail -3 $OUT/${TAG}_sanitizer_memcheck.log
