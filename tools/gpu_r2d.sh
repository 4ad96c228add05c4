#!/bin/bash
# Round 2, call D (1 GPU): new parity tests (v1.x, reference front-end on the drop-in), transform-kernel experiments.
set -u
TAG=${1:-r02d}
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests/test_hca_v1.py tests/test_reference_frontend_gpu.py -m gpu -q > $OUT/${TAG}_pytest_new.log 2>&1
tail -15 $OUT/${TAG}_pytest_new.log
for lib in base scalar pairs7; do
  for xf in 0 1; do
    CRI_LIB_PATH=$PWD/pycricodecs_b200/libcricodecs_b200_$lib.so CRI_HCA_XF=$xf timeout 300 python bench.py --no-cpu --no-companion --no-gather --e2e-steps 1 > $OUT/${TAG}_exp_${lib}_xf$xf.json 2> $OUT/${TAG}_exp_${lib}_xf$xf.err
    tail -2 $OUT/${TAG}_exp_${lib}_xf$xf.err
    python -c "
import json; d = json.load(open('$OUT/${TAG}_exp_${lib}_xf$xf.json')); print('$lib xf $xf ms', round(d['ms_per_step'], 3), [(k['kernel'], round(k['kernel_ms'], 3)) for k in d['roofline']['kernels']], d['parity_spot_check'])"
  done
done
ls -la $OUT | grep ${TAG}
