#!/bin/bash
# Round 2, call C (1 GPU): warp-pair transform kernel against the thread-resident one; crypt kernel with the group table.
set -u
TAG=${1:-r02c}
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest_gpu.log 2>&1
echo "pytest rc=$?" >> $OUT/${TAG}_pytest_gpu.log
tail -5 $OUT/${TAG}_pytest_gpu.log
CRI_HCA_XF=2 timeout 900 python -m pytest tests/test_hca_decode_gpu.py tests/test_full_size_gpu.py -m gpu -x -q > $OUT/${TAG}_pytest_xf2.log 2>&1
tail -2 $OUT/${TAG}_pytest_xf2.log
for xf in 0 1 2; do
  CRI_HCA_XF=$xf timeout 300 python bench.py --no-cpu --no-companion --e2e-steps 2 > $OUT/${TAG}_bench_xf$xf.json 2> $OUT/${TAG}_bench_xf$xf.err
  tail -2 $OUT/${TAG}_bench_xf$xf.err
  python -c "
import json; d = json.load(open('$OUT/${TAG}_bench_xf$xf.json')); print('xf $xf ms', d['ms_per_step'], [(k['kernel'], round(k['kernel_ms'], 3)) for k in d['roofline']['kernels']], d['parity_spot_check'], 'dev', d['e2e_device']['ms_per_step'], d['e2e_device']['matches_host_path'])"
done
for q in 2 3; do
  timeout 300 python bench.py --no-cpu --no-companion --e2e-steps 1 --quality $q > $OUT/${TAG}_bench_q$q.json 2> $OUT/${TAG}_bench_q$q.err
  python -c "
import json; d = json.load(open('$OUT/${TAG}_bench_q$q.json')); print('quality $q ms', d['ms_per_step'], [(k['kernel'], round(k['kernel_ms'], 3)) for k in d['roofline']['kernels']], d['parity_spot_check'])"
done
timeout 300 python bench.py --workload hca_decrypt --no-cpu --e2e-steps 2 > $OUT/${TAG}_bench_hca_decrypt.json 2> $OUT/${TAG}_bench_hca_decrypt.err
python -c "
import json; d = json.load(open('$OUT/${TAG}_bench_hca_decrypt.json')); print('decrypt', d['ms_per_step'], d['roofline']['frac'], d['parity_spot_check'], d['e2e_device']['ms_per_step'])"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:hca_imdct_pair -s 4 -c 1 -o $OUT/${TAG}_prof_imdct_pair -f \
    python bench.py --steps 1 --warmup 3 --no-cpu --no-companion --e2e-steps 1 > $OUT/${TAG}_ncu_imdct.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:hca_crypt_lut -s 4 -c 1 -o $OUT/${TAG}_prof_hca_decrypt -f \
    python bench.py --workload hca_decrypt --steps 1 --warmup 3 --no-cpu --e2e-steps 1 > $OUT/${TAG}_ncu_decrypt.log 2>&1
ls -la $OUT | grep ${TAG}
