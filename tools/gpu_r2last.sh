#!/bin/bash
# Round 2, last pass (1 GPU): everything green on the final build, default line, encoder line + capture.
set -u
TAG=${1:-r02h2}; OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/${TAG}_pytest_gpu.log; tail -3 $OUT/${TAG}_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1; tail -1 $OUT/${TAG}_smoke.log
timeout 600 python bench.py > $OUT/${TAG}_bench_hca_decode.json 2> $OUT/${TAG}_bench_hca_decode.err
timeout 600 python bench.py --workload hca_encode --cpu-seconds 5 > $OUT/${TAG}_bench_hca_encode.json 2> $OUT/${TAG}_bench_hca_encode.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:hca_encode_kernel -s 4 -c 1 -o $OUT/${TAG}_prof_hca_encode -f \
    python bench.py --workload hca_encode --steps 1 --warmup 3 --no-cpu --e2e-steps 1 > $OUT/${TAG}_ncu_encode.log 2>&1
python - <<PY
import json
for w in ("hca_decode", "hca_encode"):
    d = json.load(open("$OUT/${TAG}_bench_%s.json" % w)); r = d["roofline"]
    print(w, "value %.4g" % d["value"], "ms", round(d["ms_per_step"], 3), "frac", round(r["frac"], 4), "e2e", round(d["e2e"]["ms_per_step"], 1), "dev", round(d["e2e_device"]["ms_per_step"], 2), d["clocks"]["reasons"], d.get("parity_spot_check"))
d = json.load(open("$OUT/${TAG}_bench_hca_decode.json"))
print({k: round(d[k]["ms_per_step"], 3) for k in ("adx_encode", "hca_decrypt_decode", "hca_encode")})
PY
