#!/bin/bash
# quick GPU check: decode tests + a short default bench; usage: bash tools/gpu_quick.sh <tag> [pytest args]
TAG=${1:-q}; shift
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q "$@" > $OUT/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/${TAG}_pytest.log
tail -25 $OUT/${TAG}_pytest.log
timeout 600 python bench.py --cpu-seconds 2 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; tail -3 $OUT/${TAG}_bench.err
python - <<PY
import json
try:
    d=json.load(open("$OUT/${TAG}_bench.json")); r=d["roofline"]
    print("value", d["value"], "ms", d["ms_per_step"], "dom", r["kernel_ms"], "frac", r["frac"], "e2e", d["e2e"]["ms_per_step"], "parity", d["parity_spot_check"], d["clocks"])
except Exception as e: print("bench failed", e)
PY
