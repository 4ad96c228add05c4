#!/bin/bash
# Round 2, call AM (1 GPU): A/B of an encoder variant against the committed build in one session, with the parity tests on the variant.
set -u
V=${1:-fr}
CRI_LIB_PATH=$PWD/pycricodecs_b200/libcricodecs_b200_$V.so timeout 600 python -m pytest tests/test_hca_encode_gpu.py tests/test_full_size_gpu.py -m gpu -x -q 2>&1 | tail -2
bash tools/gpu_r2t.sh main $V main $V
