#!/bin/bash
# One gpurun call: GPU parity tests, the bench lines, the ncu launch list and full captures of the top kernels.
# Usage (from the repo root on the GPU box): bash tools/gpu_round.sh <tag> [quick]
set -u
TAG=${1:-r01}
MODE=${2:-full}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi -L > $OUT/${TAG}_gpu.txt 2>&1
nproc >> $OUT/${TAG}_gpu.txt; lscpu | grep -E 'Model name|^CPU\(s\)' >> $OUT/${TAG}_gpu.txt

timeout 900 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest_gpu.log 2>&1
echo "pytest rc=$?" >> $OUT/${TAG}_pytest_gpu.log
tail -5 $OUT/${TAG}_pytest_gpu.log

timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1
tail -2 $OUT/${TAG}_smoke.log

timeout 600 python bench.py > $OUT/${TAG}_bench_hca_decode.json 2> $OUT/${TAG}_bench_hca_decode.err
cat $OUT/${TAG}_bench_hca_decode.json
for w in adx_encode adx_decode hca_encode hca_decrypt hca_decrypt_decode; do
  timeout 600 python bench.py --workload $w --cpu-seconds 5 > $OUT/${TAG}_bench_$w.json 2> $OUT/${TAG}_bench_$w.err
  cat $OUT/${TAG}_bench_$w.json
done
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/${TAG}_bench_reference.json 2> $OUT/${TAG}_bench_reference.err
cat $OUT/${TAG}_bench_reference.json

[ "$MODE" = quick ] && exit 0

# launch list of the default bench command (our kernels only; synthesis kernels of torch are filtered out by name)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'hca_|adx_|scatter_' -c 400 --csv \
    --log-file $OUT/${TAG}_launches_hca_decode.csv python bench.py --steps 2 --warmup 3 --no-cpu --e2e-steps 1 > $OUT/${TAG}_ncu_bench.log 2>&1
# full captures of the two decode kernels (one launch each, after the warm-up launches)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:hca_imdct_fast -s 4 -c 1 -o $OUT/${TAG}_prof_imdct -f \
    python bench.py --steps 1 --warmup 3 --no-cpu --e2e-steps 1 > $OUT/${TAG}_ncu_imdct.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:hca_unpack_fast -s 4 -c 1 -o $OUT/${TAG}_prof_unpack -f \
    python bench.py --steps 1 --warmup 3 --no-cpu --e2e-steps 1 > $OUT/${TAG}_ncu_unpack.log 2>&1
for w in adx_encode adx_decode hca_encode hca_decrypt; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:"${w%%_*}_.*${w##*_}|hca_crypt" -s 4 -c 1 -o $OUT/${TAG}_prof_$w -f \
      python bench.py --workload $w --steps 1 --warmup 3 --no-cpu --e2e-steps 1 > $OUT/${TAG}_ncu_$w.log 2>&1
done
ls -la $OUT | tail -30
