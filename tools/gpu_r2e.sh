#!/bin/bash
# Round 2, call E (2 GPUs): sharded product path over NCCL (parity test), bench at N = 2, PCIe ceiling with two ranks.
set -u
TAG=${1:-r02e}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi -L > $OUT/${TAG}_gpu.txt 2>&1
nvidia-smi topo -m >> $OUT/${TAG}_gpu.txt 2>&1
lscpu | grep -E 'Model name|^CPU\(s\)|NUMA' >> $OUT/${TAG}_gpu.txt
timeout 900 python -m pytest tests/test_sharding_gpu.py tests/test_reference_frontend_gpu.py -m gpu -q > $OUT/${TAG}_pytest_sharding.log 2>&1
tail -15 $OUT/${TAG}_pytest_sharding.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --cpu-seconds 3 \
    > $OUT/${TAG}_bench_hca_decode_2gpu.json 2> $OUT/${TAG}_bench_hca_decode_2gpu.err
tail -5 $OUT/${TAG}_bench_hca_decode_2gpu.err
python - <<PY
import json
d = json.load(open("$OUT/${TAG}_bench_hca_decode_2gpu.json"))
print("value", d["value"], "ms", d["ms_per_step"], "frac", d["roofline"]["frac"])
print("e2e", d["e2e"]["ms_per_step"], "ceiling", d["e2e"]["pcie_ceiling_ms"], "numa", d["e2e"]["numa"], "dev", d["e2e_device"]["ms_per_step"], d["e2e_device"]["matches_host_path"])
print("gather", d.get("gather"))
for k in ("adx_encode", "hca_decrypt_decode", "hca_encode"):
    print(k, d.get(k))
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --workload hca_encode --scaling strong --streams 4096 --no-cpu \
    > $OUT/${TAG}_bench_hca_encode_strong_2gpu.json 2> $OUT/${TAG}_bench_hca_encode_strong_2gpu.err
tail -3 $OUT/${TAG}_bench_hca_encode_strong_2gpu.err
python -c "
import json; d = json.load(open('$OUT/${TAG}_bench_hca_encode_strong_2gpu.json')); print('encode strong 2gpu', d['value'], d['ms_per_step'], d['scaling'], d.get('gather'))"
bash tools/pcie_ceiling.sh 2 > $OUT/${TAG}_pcie_ceiling_2gpu.json 2>&1
tail -1 $OUT/${TAG}_pcie_ceiling_2gpu.json
ls -la $OUT | grep ${TAG}
