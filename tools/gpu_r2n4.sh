#!/bin/bash
# Round 2 (4 GPUs): the default bench line at N = 4 (kernels, e2e against the 4-rank PCIe ceiling, companions, gather).
set -u
TAG=${1:-r02bc}; OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus 4 --no-cpu \
    > $OUT/${TAG}_bench_hca_decode_4gpu.json 2> $OUT/${TAG}_bench_hca_decode_4gpu.err
tail -2 $OUT/${TAG}_bench_hca_decode_4gpu.err | cut -c1-200
python - <<PY
import json
d = json.load(open("$OUT/${TAG}_bench_hca_decode_4gpu.json"))
print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["ms_per_step"], "ceiling", d["e2e"]["pcie_ceiling_ms"], "dev", d["e2e_device"]["ms_per_step"])
print("gather", {k: round(v["ms_per_step"], 2) for k, v in d["gather"].items() if isinstance(v, dict)})
for k in ("adx_encode", "hca_decrypt_decode", "hca_encode"):
    print(k, round(d[k]["ms_per_step"], 3), d[k]["value"])
PY
