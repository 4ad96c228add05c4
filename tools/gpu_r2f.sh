#!/bin/bash
# Round 2, call F (1 GPU): pipelined HCA decode (unpack of chunk k+1 beside the transform of chunk k): parity + sweep;
# regression tests of the review fixes.
set -u
TAG=${1:-r02f}
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest_gpu.log 2>&1
tail -4 $OUT/${TAG}_pytest_gpu.log
CRI_HCA_CHUNKS=4 timeout 900 python -m pytest tests/test_hca_decode_gpu.py tests/test_full_size_gpu.py tests/test_hca_v1.py tests/test_batch_pipeline_gpu.py -m gpu -x -q > $OUT/${TAG}_pytest_chunks4.log 2>&1
tail -4 $OUT/${TAG}_pytest_chunks4.log
for k in 1 2 4 6 8 12 16; do
  CRI_HCA_CHUNKS=$k timeout 300 python bench.py --no-cpu --no-companion --no-gather --e2e-steps 1 > $OUT/${TAG}_bench_chunks$k.json 2> $OUT/${TAG}_bench_chunks$k.err
  tail -2 $OUT/${TAG}_bench_chunks$k.err
  python -c "
import json; d = json.load(open('$OUT/${TAG}_bench_chunks$k.json')); print('chunks $k ms', round(d['ms_per_step'], 3), 'launches', d['gpu_launches'], d['parity_spot_check'], 'dev', round(d['e2e_device']['ms_per_step'], 2))"
done
for q in 2 3; do
  CRI_HCA_CHUNKS=4 timeout 300 python bench.py --no-cpu --no-companion --no-gather --e2e-steps 1 --quality $q > $OUT/${TAG}_bench_chunks4_q$q.json 2> $OUT/${TAG}_bench_chunks4_q$q.err
  python -c "
import json; d = json.load(open('$OUT/${TAG}_bench_chunks4_q$q.json')); print('chunks 4 quality $q ms', round(d['ms_per_step'], 3), d['parity_spot_check'])"
done
ls -la $OUT | grep ${TAG} | head -40
