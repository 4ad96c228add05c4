#!/bin/bash
# Round 2, call AP (1 GPU): tests that use the HCA class surface (after its internals went from streams to byte strings).
timeout 900 python -m pytest tests/test_hca_encode_gpu.py tests/test_hca_crypt_gpu.py tests/test_hca_decode_gpu.py tests/test_dropin_gpu.py tests/test_reference_frontend_gpu.py tests/test_usm_audio.py tests/test_awb.py tests/test_acb.py -m gpu -x -q 2>&1 | tail -4
