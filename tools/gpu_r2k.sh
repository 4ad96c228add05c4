#!/bin/bash
# Round 2, call K (1 GPU): everything once more after the host-planning and ADX-mover changes; sanitizer over the new kernels.
set -u
TAG=${1:-r02k}
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest_gpu.log 2>&1
tail -4 $OUT/${TAG}_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1
tail -2 $OUT/${TAG}_smoke.log
timeout 600 python bench.py --cpu-seconds 5 > $OUT/${TAG}_bench_hca_decode.json 2> $OUT/${TAG}_bench_hca_decode.err
tail -2 $OUT/${TAG}_bench_hca_decode.err
python - <<PY
import json
d = json.load(open("$OUT/${TAG}_bench_hca_decode.json"))
print("value", d["value"], "ms", d["ms_per_step"], "frac", d["roofline"]["frac"], [(k["kernel"], round(k["kernel_ms"], 3)) for k in d["roofline"]["kernels"]])
print("e2e", d["e2e"]["ms_per_step"], "ceiling", d["e2e"]["pcie_ceiling_ms"], "dev", d["e2e_device"]["ms_per_step"], d["e2e_device"]["matches_host_path"])
for k in ("adx_encode", "hca_decrypt_decode", "hca_encode"):
    print(k, d[k]["ms_per_step"], d[k]["roofline_frac"])
print(d["cpu_baseline"])
PY
for w in adx_encode adx_decode hca_encode hca_decrypt; do
  timeout 300 python bench.py --workload $w --no-cpu --e2e-steps 2 > $OUT/${TAG}_bench_$w.json 2> $OUT/${TAG}_bench_$w.err
  python -c "
import json; d = json.load(open('$OUT/${TAG}_bench_$w.json')); print('$w ms', round(d['ms_per_step'], 3), 'frac', round(d['roofline']['frac'], 4), d['parity_spot_check'], 'e2e', round(d['e2e']['ms_per_step'], 1), 'dev', round(d['e2e_device']['ms_per_step'], 2), d['e2e_device']['matches_host_path'])"
done
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 1 python -m pytest tests/test_hca_encode_gpu.py tests/test_adx_gpu.py tests/test_hca_crypt_gpu.py tests/test_regressions_gpu.py -m gpu -x -q > $OUT/${TAG}_sanitizer_memcheck.log 2>&1
echo "memcheck rc=$?" >> $OUT/${TAG}_sanitizer_memcheck.log
tail -3 $OUT/${TAG}_sanitizer_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 1 python -m pytest tests/test_hca_encode_gpu.py tests/test_adx_gpu.py -m gpu -x -q -k "not looping" > $OUT/${TAG}_sanitizer_racecheck.log 2>&1
echo "racecheck rc=$?" >> $OUT/${TAG}_sanitizer_racecheck.log
tail -3 $OUT/${TAG}_sanitizer_racecheck.log
ls -la $OUT | grep ${TAG} | wc -l
