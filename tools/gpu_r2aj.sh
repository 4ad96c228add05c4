#!/bin/bash
# Round 2, call AJ (2 GPUs): timeline of the sharded path's chunks (when each chunk's compute ends, when its all-gather runs).
set -u
CRI_GATHER_TRACE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 2 --no-cpu --no-companion --steps 3 --warmup 3 --e2e-steps 1 2>&1 >/dev/null | grep "gather trace" | cut -c1-700
