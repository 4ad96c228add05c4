#!/bin/bash
# Round 2, call AC (1 GPU): host-side phases of the device-pointer batch calls (CRI_TRACE=1).
set -u
OUT=gpurun_out
for w in hca_decode adx_encode hca_encode adx_decode hca_decrypt; do
  echo "== $w"
  CRI_TRACE=1 timeout 300 python bench.py --workload $w --no-cpu --no-companion --steps 3 --warmup 3 --e2e-steps 2 2>&1 >/dev/null | grep "cri trace" | tail -3
done
