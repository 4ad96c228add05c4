#!/bin/bash
# Round 2, call G (1 GPU): HCA encode kernel after the bit-cost / PCM-load / two-phase packing rework.
set -u
TAG=${1:-r02g}
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest_gpu.log 2>&1
tail -6 $OUT/${TAG}_pytest_gpu.log
timeout 300 python bench.py --workload hca_encode --no-cpu --e2e-steps 2 > $OUT/${TAG}_bench_hca_encode.json 2> $OUT/${TAG}_bench_hca_encode.err
tail -2 $OUT/${TAG}_bench_hca_encode.err
python -c "
import json; d = json.load(open('$OUT/${TAG}_bench_hca_encode.json')); print('hca_encode ms', round(d['ms_per_step'], 3), 'frac', round(d['roofline']['frac'], 4), d['parity_spot_check'], 'dev', round(d['e2e_device']['ms_per_step'], 2))"
for q in 0 2 3; do
timeout 300 python bench.py --workload hca_encode --no-cpu --e2e-steps 1 --quality $q > $OUT/${TAG}_bench_hca_encode_q$q.json 2> $OUT/${TAG}_bench_hca_encode_q$q.err
python -c "
import json; d = json.load(open('$OUT/${TAG}_bench_hca_encode_q$q.json')); print('hca_encode quality $q ms', round(d['ms_per_step'], 3), d['parity_spot_check'])"
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:hca_encode_kernel -s 4 -c 1 -o $OUT/${TAG}_prof_hca_encode -f \
    python bench.py --workload hca_encode --steps 1 --warmup 3 --no-cpu --e2e-steps 1 > $OUT/${TAG}_ncu_encode.log 2>&1
ls -la $OUT | grep ${TAG}
