#!/bin/bash
# Round 2, call B (1 GPU): parity, default bench line (new layout), crypt kernel v2 + capture.
set -u
TAG=${1:-r02b}
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest_gpu.log 2>&1
echo "pytest rc=$?" >> $OUT/${TAG}_pytest_gpu.log
tail -5 $OUT/${TAG}_pytest_gpu.log
timeout 600 python bench.py --cpu-seconds 5 > $OUT/${TAG}_bench_hca_decode.json 2> $OUT/${TAG}_bench_hca_decode.err
tail -3 $OUT/${TAG}_bench_hca_decode.err
python - <<PY
import json
d = json.load(open("$OUT/${TAG}_bench_hca_decode.json"))
print("value", d["value"], "ms", d["ms_per_step"], "wall", d["wall_ms_per_step"], "frac", d["roofline"]["frac"])
print("e2e", d["e2e"]["ms_per_step"], "ceiling", d["e2e"]["pcie_ceiling_ms"], "dev", d["e2e_device"])
for k in ("adx_encode", "hca_decrypt_decode", "hca_encode"):
    print(k, d.get(k))
PY
timeout 300 python bench.py --workload hca_decrypt --no-cpu --e2e-steps 2 > $OUT/${TAG}_bench_hca_decrypt.json 2> $OUT/${TAG}_bench_hca_decrypt.err
tail -3 $OUT/${TAG}_bench_hca_decrypt.err
python -c "
import json; d = json.load(open('$OUT/${TAG}_bench_hca_decrypt.json')); print('decrypt', d['ms_per_step'], d['roofline']['frac'], d['parity_spot_check'], d['e2e_device'])"
timeout 300 python bench.py --workload hca_encode --scaling strong --streams 4096 --no-cpu --e2e-steps 2 > $OUT/${TAG}_bench_hca_encode_strong.json 2> $OUT/${TAG}_bench_hca_encode_strong.err
tail -3 $OUT/${TAG}_bench_hca_encode_strong.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:hca_crypt_lut -s 4 -c 1 -o $OUT/${TAG}_prof_hca_decrypt -f \
    python bench.py --workload hca_decrypt --steps 1 --warmup 3 --no-cpu --e2e-steps 1 > $OUT/${TAG}_ncu_decrypt.log 2>&1
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/${TAG}_bench_reference.json 2> $OUT/${TAG}_bench_reference.err
cat $OUT/${TAG}_bench_reference.json | cut -c1-400
ls -la $OUT | grep ${TAG}
