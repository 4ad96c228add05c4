#!/bin/bash
# Round 2, call J (8 GPUs): the bench lines the driver's scaling run records, at N = 8: default (configs[1] + companions +
# gather), configs[3] (65536 encrypted streams over 8 GPUs), configs[4] (4096-stream encode, strong), PCIe ceiling x8.
set -u
TAG=${1:-r02j}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi -L > $OUT/${TAG}_gpu.txt 2>&1
nvidia-smi topo -m >> $OUT/${TAG}_gpu.txt 2>&1
lscpu | grep -E 'Model name|^CPU\(s\)|NUMA' >> $OUT/${TAG}_gpu.txt
run() { # name, extra args
  local name=$1; shift
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 "$@" \
      > $OUT/${TAG}_bench_${name}_8gpu.json 2> $OUT/${TAG}_bench_${name}_8gpu.err
  tail -3 $OUT/${TAG}_bench_${name}_8gpu.err | cut -c1-300
}
run hca_decode --cpu-seconds 3
python - <<PY
import json
d = json.load(open("$OUT/${TAG}_bench_hca_decode_8gpu.json"))
print("value", d["value"], "ms", d["ms_per_step"], "frac", d["roofline"]["frac"])
print("e2e", d["e2e"]["ms_per_step"], "ceiling", d["e2e"]["pcie_ceiling_ms"], "numa", d["e2e"]["numa"], "dev", d["e2e_device"]["ms_per_step"], d["e2e_device"]["matches_host_path"])
print("gather", d.get("gather"))
for k in ("adx_encode", "hca_decrypt_decode", "hca_encode"):
    print(k, d.get(k))
PY
run hca_decrypt_decode --workload hca_decrypt_decode --no-cpu --no-companion
python -c "
import json; d = json.load(open('$OUT/${TAG}_bench_hca_decrypt_decode_8gpu.json')); print('decrypt+decode 8gpu', d['value'], d['ms_per_step'], d['config']['streams_per_gpu'], 'gather', d.get('gather'))"
run hca_encode_strong --workload hca_encode --scaling strong --streams 4096 --no-cpu
python -c "
import json; d = json.load(open('$OUT/${TAG}_bench_hca_encode_strong_8gpu.json')); print('encode strong 8gpu', d['value'], d['ms_per_step'], d['scaling'], d['config']['streams_per_gpu'], 'gather', d.get('gather'))"
bash tools/pcie_ceiling.sh 8 > $OUT/${TAG}_pcie_ceiling_8gpu.json 2>&1
tail -1 $OUT/${TAG}_pcie_ceiling_8gpu.json
bash tools/pcie_ceiling.sh 4 > $OUT/${TAG}_pcie_ceiling_4gpu.json 2>&1
tail -1 $OUT/${TAG}_pcie_ceiling_4gpu.json
ls -la $OUT | grep ${TAG}
