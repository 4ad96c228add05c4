#!/bin/bash
# Round 2, call A (1 GPU): parity (plain and with poisoned output blobs), the new crypt kernel against the staged one,
# bare PCIe ceiling, e2e chunk sweep, sanitizer over the new entry points.
set -u
TAG=${1:-r02a}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi -L > $OUT/${TAG}_gpu.txt 2>&1
nproc >> $OUT/${TAG}_gpu.txt; lscpu | grep -E 'Model name|^CPU\(s\)|NUMA' >> $OUT/${TAG}_gpu.txt
nvidia-smi topo -m >> $OUT/${TAG}_gpu.txt 2>&1

timeout 900 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest_gpu.log 2>&1
echo "pytest rc=$?" >> $OUT/${TAG}_pytest_gpu.log
tail -5 $OUT/${TAG}_pytest_gpu.log
CRI_POISON=1 timeout 900 python -m pytest tests -m gpu -q > $OUT/${TAG}_pytest_poison.log 2>&1
echo "pytest rc=$?" >> $OUT/${TAG}_pytest_poison.log
tail -15 $OUT/${TAG}_pytest_poison.log

CRI_HCA_CRYPT_STAGED=1 timeout 300 python bench.py --workload hca_decrypt --no-cpu --e2e-steps 2 > $OUT/${TAG}_bench_hca_decrypt_staged.json 2> $OUT/${TAG}_bench_hca_decrypt_staged.err
timeout 300 python bench.py --workload hca_decrypt --no-cpu --e2e-steps 2 > $OUT/${TAG}_bench_hca_decrypt.json 2> $OUT/${TAG}_bench_hca_decrypt.err
cat $OUT/${TAG}_bench_hca_decrypt_staged.json $OUT/${TAG}_bench_hca_decrypt.json | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('decrypt', d['ms_per_step'], d['roofline']['frac'], d['parity_spot_check'])"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:hca_crypt_lut -s 4 -c 1 -o $OUT/${TAG}_prof_hca_decrypt -f \
    python bench.py --workload hca_decrypt --steps 1 --warmup 3 --no-cpu --e2e-steps 1 > $OUT/${TAG}_ncu_decrypt.log 2>&1

bash tools/pcie_ceiling.sh 1 > $OUT/${TAG}_pcie_ceiling_1gpu.json 2>&1
bash tools/pcie_ceiling.sh 1 --numa > $OUT/${TAG}_pcie_ceiling_1gpu_numa.json 2>&1
cat $OUT/${TAG}_pcie_ceiling_1gpu.json $OUT/${TAG}_pcie_ceiling_1gpu_numa.json

for mb in 64 128 256 512; do
  CRI_TRACE=1 CRI_CHUNK_MB=$mb timeout 300 python bench.py --no-cpu --no-companion --steps 3 --e2e-steps 4 > $OUT/${TAG}_e2e_chunk${mb}.json 2> $OUT/${TAG}_e2e_chunk${mb}.err
  python -c "
import json; d = json.load(open('$OUT/${TAG}_e2e_chunk${mb}.json')); print('chunk $mb MB e2e ms', d['e2e']['ms_per_step'], 'kernels ms', d['ms_per_step'])"
done
timeout 300 python bench.py --no-cpu --no-companion --steps 3 --e2e-steps 4 > $OUT/${TAG}_e2e_default.json 2> $OUT/${TAG}_e2e_default.err
python -c "
import json; d = json.load(open('$OUT/${TAG}_e2e_default.json')); print('default e2e ms', d['e2e']['ms_per_step'], 'kernels ms', d['ms_per_step'])"

timeout 900 compute-sanitizer --tool memcheck --error-exitcode 1 python -m pytest tests/test_device_api_gpu.py tests/test_hca_crypt_gpu.py -m gpu -x -q > $OUT/${TAG}_sanitizer_memcheck.log 2>&1
echo "memcheck rc=$?" >> $OUT/${TAG}_sanitizer_memcheck.log
tail -4 $OUT/${TAG}_sanitizer_memcheck.log
ls -la $OUT | tail -30
