#!/bin/bash
# Round 2, call L (1 GPU): kept-intensity decode (scan kernel), smoke, racecheck of the crypt kernel, joint-stereo decode timing.
set -u
TAG=${1:-r02l}
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest_gpu.log 2>&1
tail -15 $OUT/${TAG}_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1
tail -2 $OUT/${TAG}_smoke.log
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 1 python -m pytest tests/test_hca_crypt_gpu.py tests/test_hca_v3.py -m gpu -x -q > $OUT/${TAG}_sanitizer_racecheck.log 2>&1
echo "racecheck rc=$?" >> $OUT/${TAG}_sanitizer_racecheck.log
tail -3 $OUT/${TAG}_sanitizer_racecheck.log
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 1 python -m pytest tests/test_hca_v3.py tests/test_hca_decode_gpu.py -m gpu -x -q > $OUT/${TAG}_sanitizer_memcheck.log 2>&1
echo "memcheck rc=$?" >> $OUT/${TAG}_sanitizer_memcheck.log
tail -3 $OUT/${TAG}_sanitizer_memcheck.log
for q in 1 3; do
  timeout 300 python bench.py --workload hca_decode --quality $q --no-cpu --e2e-steps 2 > $OUT/${TAG}_bench_hca_decode_q$q.json 2> $OUT/${TAG}_bench_hca_decode_q$q.err
  tail -2 $OUT/${TAG}_bench_hca_decode_q$q.err
  python -c "
import json; d = json.load(open('$OUT/${TAG}_bench_hca_decode_q$q.json')); print('q$q ms', round(d['ms_per_step'], 3), 'launches', d['gpu_launches'], [(k['kernel'], round(k['kernel_ms'], 3)) for k in d['roofline']['kernels']], d['parity_spot_check'])"
done
