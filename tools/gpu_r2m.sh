#!/bin/bash
# Round 2, call M (1 GPU): unpack kernel changes (phase-1 ring, prefix-codebook variant): decode tests, bench, ncu of the unpack kernel.
set -u
TAG=${1:-r02m}
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests/test_hca_decode_gpu.py tests/test_hca_v3.py tests/test_hca_v1.py tests/test_full_size_gpu.py tests/test_regressions_gpu.py -m gpu -x -q > $OUT/${TAG}_pytest_gpu.log 2>&1
tail -8 $OUT/${TAG}_pytest_gpu.log
for q in 1 3; do
  timeout 300 python bench.py --workload hca_decode --quality $q --no-cpu --e2e-steps 2 > $OUT/${TAG}_bench_hca_decode_q$q.json 2> $OUT/${TAG}_bench_hca_decode_q$q.err
  tail -2 $OUT/${TAG}_bench_hca_decode_q$q.err
  python -c "
import json; d = json.load(open('$OUT/${TAG}_bench_hca_decode_q$q.json')); print('q$q ms', round(d['ms_per_step'], 3), [(k['kernel'], round(k['kernel_ms'], 3)) for k in d['roofline']['kernels']], d['parity_spot_check'])"
done
if [ "${2:-}" = "ncu" ]; then
  timeout 600 ncu --set full --import-source on --clock-control none -k regex:hca_unpack_fast -c 1 -s 2 -o $OUT/${TAG}_prof_unpack python bench.py --workload hca_decode --no-cpu --steps 1 --warmup 1 --e2e-steps 0 > $OUT/${TAG}_ncu_unpack.log 2>&1
  tail -2 $OUT/${TAG}_ncu_unpack.log
fi
