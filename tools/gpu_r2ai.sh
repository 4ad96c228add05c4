#!/bin/bash
# Round 2, call AI (1 GPU): encoder with the CTA-level stream lookup: parity, then multi-round grids (CRI_LIB_PATH variants).
set -u
timeout 600 python -m pytest tests/test_hca_encode_gpu.py tests/test_wav_ingest.py tests/test_full_size_gpu.py tests/test_usm_audio.py -m gpu -x -q 2>&1 | tail -2
bash tools/gpu_r2t.sh main encr2 encr3 encr4 encr8 main
