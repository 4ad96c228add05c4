#!/bin/bash
# Round 2, call AN (1 GPU): the encoder's hard-material test.
timeout 600 python -m pytest tests/test_hca_encode_gpu.py -m gpu -x -q 2>&1 | tail -8
