"""ctypes loader of libcricodecs_b200.so (the C-ABI in include/cricodecs_b200.h).

The library is built in-tree by pycricodecs_b200/csrc/Makefile (see
__graft_entry__.build). There is no fallback: if the shared object is missing
the import of any compute entry point raises, and without a CUDA device
`Context()` raises -- nothing here ever routes to a CPU implementation.
"""
from __future__ import annotations

import ctypes
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("CRI_LIB_PATH") or os.path.join(HERE, "libcricodecs_b200.so")   # CRI_LIB_PATH: kernel experiments (tools/build_variant.sh)

c_u8p = ctypes.POINTER(ctypes.c_uint8)
c_u64p = ctypes.POINTER(ctypes.c_uint64)
c_u16p = ctypes.POINTER(ctypes.c_uint16)
c_i32p = ctypes.POINTER(ctypes.c_int32)
c_vp = ctypes.c_void_p


class AdxParams(ctypes.Structure):
    _fields_ = [("bit_depth", ctypes.c_uint32), ("block_size", ctypes.c_uint32), ("encoding", ctypes.c_uint32),
                ("highpass", ctypes.c_uint32), ("filter", ctypes.c_uint32), ("version", ctypes.c_uint32),
                ("force_not_looping", ctypes.c_uint32)]


class JobDesc(ctypes.Structure):
    _fields_ = [("kind", ctypes.c_int), ("blob", c_vp), ("offsets", c_vp), ("n", ctypes.c_uint32), ("keys", c_vp),
                ("subkeys", c_vp), ("adx", AdxParams), ("quality", ctypes.c_uint32), ("encrypt", ctypes.c_int),
                ("ciph_type", ctypes.c_uint32), ("d_blob", c_vp), ("d_out", c_vp), ("stream", c_vp)]


JOB_ADX_DECODE, JOB_ADX_ENCODE, JOB_HCA_DECODE, JOB_HCA_CRYPT, JOB_HCA_ENCODE = 1, 2, 3, 4, 5

# name -> (restype, argtypes); also the list tests check against include/cricodecs_b200.h
SIGNATURES = {
    "cri_ctx_create": (ctypes.c_int, [ctypes.c_int, ctypes.POINTER(c_vp)]),
    "cri_ctx_destroy": (None, [c_vp]),
    "cri_ctx_trim": (None, [c_vp]),
    "cri_host_alloc": (c_vp, [ctypes.c_size_t]),
    "cri_host_free": (None, [c_vp]),
    "cri_last_error": (ctypes.c_char_p, [c_vp]),
    "cri_version": (ctypes.c_int, []),
    "cri_ctx_launch_count": (ctypes.c_uint64, [c_vp]),
    "cri_ctx_last_kernel_ms": (ctypes.c_float, [c_vp]),
    "cri_ctx_last_dominant_ms": (ctypes.c_float, [c_vp]),
    "cri_adx_decode_sizes": (ctypes.c_int, [c_vp, c_vp, ctypes.c_uint32, c_vp, c_vp]),
    "cri_adx_decode_batch": (ctypes.c_int, [c_vp, c_vp, c_vp, ctypes.c_uint32, c_vp, c_vp, c_vp]),
    "cri_adx_encode_sizes": (ctypes.c_int, [c_vp, c_vp, ctypes.c_uint32, ctypes.POINTER(AdxParams), c_vp, c_vp]),
    "cri_adx_encode_batch": (ctypes.c_int, [c_vp, c_vp, c_vp, ctypes.c_uint32, ctypes.POINTER(AdxParams), c_vp, c_vp, c_vp]),
    "cri_hca_decode_sizes": (ctypes.c_int, [c_vp, c_vp, ctypes.c_uint32, c_vp, c_vp]),
    "cri_hca_decode_batch": (ctypes.c_int, [c_vp, c_vp, c_vp, ctypes.c_uint32, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "cri_hca_crypt_batch": (ctypes.c_int, [c_vp, c_vp, c_vp, ctypes.c_uint32, ctypes.c_int, ctypes.c_uint32, c_vp, c_vp, c_vp, c_vp]),
    "cri_hca_encode_sizes": (ctypes.c_int, [c_vp, c_vp, ctypes.c_uint32, ctypes.c_uint32, c_vp, c_vp]),
    "cri_hca_encode_sizes_ex": (ctypes.c_int, [c_vp, c_vp, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32, c_vp, c_vp]),
    "cri_hca_encode_batch": (ctypes.c_int, [c_vp, c_vp, c_vp, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32, c_vp, c_vp, c_vp]),
    "cri_sizes_dev": (ctypes.c_int, [c_vp, ctypes.c_int, c_vp, c_vp, ctypes.c_uint32, ctypes.POINTER(AdxParams), ctypes.c_uint32, c_vp, c_vp, c_vp]),
    "cri_adx_decode_batch_dev": (ctypes.c_int, [c_vp, c_vp, c_vp, ctypes.c_uint32, c_vp, c_vp, c_vp, c_vp]),
    "cri_adx_encode_batch_dev": (ctypes.c_int, [c_vp, c_vp, c_vp, ctypes.c_uint32, ctypes.POINTER(AdxParams), c_vp, c_vp, c_vp, c_vp]),
    "cri_hca_decode_batch_dev": (ctypes.c_int, [c_vp, c_vp, c_vp, ctypes.c_uint32, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "cri_hca_crypt_batch_dev": (ctypes.c_int, [c_vp, c_vp, c_vp, ctypes.c_uint32, ctypes.c_int, ctypes.c_uint32, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "cri_hca_encode_batch_dev": (ctypes.c_int, [c_vp, c_vp, c_vp, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32, c_vp, c_vp, c_vp, c_vp]),
    "cri_job_create": (ctypes.c_int, [c_vp, ctypes.POINTER(JobDesc), ctypes.POINTER(c_vp)]),
    "cri_job_out_bytes": (ctypes.c_uint64, [c_vp]),
    "cri_job_out_offsets": (c_u64p, [c_vp]),
    "cri_job_units": (ctypes.c_uint64, [c_vp]),
    "cri_job_upload": (ctypes.c_int, [c_vp, c_vp]),
    "cri_job_run": (ctypes.c_int, [c_vp, c_vp]),
    "cri_job_download": (ctypes.c_int, [c_vp, c_vp, c_vp, c_vp]),
    "cri_job_destroy": (None, [c_vp, c_vp]),
    "cri_adx_decode": (ctypes.c_int, [c_vp, c_vp, ctypes.c_size_t, ctypes.POINTER(c_vp), ctypes.POINTER(ctypes.c_size_t)]),
    "cri_adx_encode": (ctypes.c_int, [c_vp, c_vp, ctypes.c_size_t, ctypes.POINTER(AdxParams), ctypes.POINTER(c_vp), ctypes.POINTER(ctypes.c_size_t)]),
    "cri_hca_decode": (ctypes.c_int, [c_vp, c_vp, ctypes.c_size_t, ctypes.c_uint64, ctypes.c_uint16, ctypes.POINTER(c_vp), ctypes.POINTER(ctypes.c_size_t)]),
    "cri_hca_crypt": (ctypes.c_int, [c_vp, c_vp, ctypes.c_size_t, ctypes.c_int, ctypes.c_uint32, ctypes.c_uint64, ctypes.c_uint16, ctypes.POINTER(c_vp), ctypes.POINTER(ctypes.c_size_t)]),
    "cri_hca_encode": (ctypes.c_int, [c_vp, c_vp, ctypes.c_size_t, ctypes.c_uint32, ctypes.c_uint32, ctypes.POINTER(c_vp), ctypes.POINTER(ctypes.c_size_t)]),
    "cri_free": (None, [c_vp]),
    "cri_crc16": (ctypes.c_uint16, [c_vp, ctypes.c_size_t]),
    "cri_hca_cipher_table": (ctypes.c_int, [ctypes.c_int, ctypes.c_uint64, c_vp]),
    "cri_hca_mix_subkey": (ctypes.c_uint64, [ctypes.c_uint64, ctypes.c_uint16]),
    "cri_adx_coefficients": (None, [ctypes.c_uint32, ctypes.c_uint32, c_vp]),
    "cri_strerror": (ctypes.c_char_p, [ctypes.c_int]),
}

_lib = None


def lib() -> ctypes.CDLL:
    """Load the shared object once; raise (never fall back) when it is absent."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} is not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
                              "(or make -C pycricodecs_b200/csrc). cricodecs_b200 has no CPU fallback.")
        L = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib
