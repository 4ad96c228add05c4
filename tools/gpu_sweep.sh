#!/bin/bash
# bench the default workload under several environment settings; usage: bash tools/gpu_sweep.sh <tag> "VAR=1 VAR2=2" "..." ...
TAG=$1; shift
OUT=gpurun_out; mkdir -p $OUT
i=0
for E in "$@"; do
  i=$((i+1))
  env $E timeout 600 python bench.py --no-cpu --e2e-steps 1 > $OUT/${TAG}_sweep_$i.json 2> $OUT/${TAG}_sweep_$i.err
  python - "$E" $OUT/${TAG}_sweep_$i.json <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[2])); r=d["roofline"]
    print(f"{sys.argv[1]:40s} total {d['ms_per_step']:.3f} ms  dominant {r['kernel_ms']:.3f} ms  parity {d['parity_spot_check']}")
except Exception as e: print(sys.argv[1], "failed", e)
PY
done
