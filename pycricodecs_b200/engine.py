"""Batch front-end over the C-ABI: many independent streams per call.

`Context` owns one CUDA stream on one GPU; `Job` keeps a batch resident in HBM
(what bench.py times); the `*_batch` helpers take a list of `bytes` and return a
list of `bytes` (or the reference's exception per failed stream).

Error mapping follows the reference: ADX / WAV errors raise ValueError (or
NotImplementedError for encrypted ADX) with the reference's message
(CriCodecs/adx.cpp:11-38, pcm.cpp:22-38), HCA errors the four messages of
py_codec_err (CriCodecs/hca.cpp:3252-3268).
"""
from __future__ import annotations

import ctypes
from typing import Iterable, List, Optional, Sequence

import numpy as np

from . import _lib
from ._lib import AdxParams, JobDesc

ERR_CUDA = -400


class CriError(ValueError):
    def __init__(self, status: int, message: str):
        super().__init__(message)
        self.status = status


def strerror(status: int) -> str:
    return _lib.lib().cri_strerror(status).decode()


def exception_for(status: int) -> Exception:
    msg = strerror(status)
    if status == -3:
        return NotImplementedError(msg)  # encrypted ADX, adx.cpp:34-35
    if status == ERR_CUDA:
        return RuntimeError(msg)
    return CriError(status, msg)


def pack(streams: Sequence[bytes]):
    """Concatenate streams into (blob uint8[total], offsets uint64[n+1])."""
    n = len(streams)
    offsets = np.zeros(n + 1, dtype=np.uint64)
    if n:
        offsets[1:] = np.cumsum([len(s) for s in streams], dtype=np.uint64)
    blob = np.frombuffer(b"".join(bytes(s) for s in streams), dtype=np.uint8) if n else np.zeros(0, np.uint8)
    return blob, offsets


class Context:
    """One GPU, one CUDA stream. Raises RuntimeError when no CUDA device is usable."""

    def __init__(self, device: int = 0):
        self._lib = _lib.lib()
        h = ctypes.c_void_p()
        rc = self._lib.cri_ctx_create(device, ctypes.byref(h))
        if rc != 0:
            raise RuntimeError(f"cricodecs_b200: cannot open CUDA device {device} (status {rc}); there is no CPU fallback")
        self.handle = h
        self.device = device

    def close(self):
        if getattr(self, "handle", None):
            self._lib.cri_ctx_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def launches(self) -> int:
        return int(self._lib.cri_ctx_launch_count(self.handle))

    @property
    def last_kernel_ms(self) -> float:
        return float(self._lib.cri_ctx_last_kernel_ms(self.handle))

    @property
    def last_dominant_ms(self) -> float:
        return float(self._lib.cri_ctx_last_dominant_ms(self.handle))

    def check(self, rc: int):
        if rc == ERR_CUDA:
            raise RuntimeError("cricodecs_b200 CUDA failure: " + self._lib.cri_last_error(self.handle).decode())
        if rc != 0:
            raise exception_for(rc)


_default_ctx: Optional[Context] = None


def default_context() -> Context:
    global _default_ctx
    if _default_ctx is None:
        _default_ctx = Context(0)
    return _default_ctx


class Job:
    """A batch kept resident in HBM: create (parse + plan + H2D), run (kernels), download (D2H)."""

    def __init__(self, ctx: Context, kind: int, blob: np.ndarray, offsets: np.ndarray, *, keys=None, subkeys=None,
                 adx: Optional[AdxParams] = None, quality: int = 1, encrypt: int = 0, ciph_type: int = 0):
        self.ctx = ctx
        self._lib = ctx._lib
        self.blob = np.ascontiguousarray(blob, dtype=np.uint8)       # must outlive the job (borrowed by the library)
        self.offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        self.n = len(self.offsets) - 1
        self._keys = None if keys is None else np.ascontiguousarray(keys, dtype=np.uint64)
        self._subkeys = None if subkeys is None else np.ascontiguousarray(subkeys, dtype=np.uint16)
        d = JobDesc()
        d.kind = kind
        d.blob = self.blob.ctypes.data
        d.offsets = self.offsets.ctypes.data
        d.n = self.n
        d.keys = None if self._keys is None else self._keys.ctypes.data
        d.subkeys = None if self._subkeys is None else self._subkeys.ctypes.data
        if adx is not None:
            d.adx = adx
        d.quality = quality
        d.encrypt = encrypt
        d.ciph_type = ciph_type
        h = ctypes.c_void_p()
        ctx.check(self._lib.cri_job_create(ctx.handle, ctypes.byref(d), ctypes.byref(h)))
        self.handle = h
        self.out_bytes = int(self._lib.cri_job_out_bytes(h))
        p = self._lib.cri_job_out_offsets(h)
        self.out_offsets = np.ctypeslib.as_array(p, shape=(self.n + 1,)).copy()
        self.units = int(self._lib.cri_job_units(h))

    def upload(self):
        self.ctx.check(self._lib.cri_job_upload(self.ctx.handle, self.handle))

    def run(self):
        self.ctx.check(self._lib.cri_job_run(self.ctx.handle, self.handle))

    def download(self, out: Optional[np.ndarray] = None):
        if out is None:
            out = np.empty(self.out_bytes, dtype=np.uint8)
        status = np.zeros(max(self.n, 1), dtype=np.int32)
        self.ctx.check(self._lib.cri_job_download(self.ctx.handle, self.handle, out.ctypes.data, status.ctypes.data))
        return out, status[: self.n]

    def close(self):
        if getattr(self, "handle", None):
            self._lib.cri_job_destroy(self.ctx.handle, self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()


def batch_blob(kind: int, blob: np.ndarray, offsets: np.ndarray, ctx: Optional[Context] = None, out: Optional[np.ndarray] = None,
               *, keys=None, subkeys=None, adx: Optional[AdxParams] = None, quality: int = 1, encrypt: int = 0,
               ciph_type: int = 0):
    """One call of the C-ABI batch entry point for `kind`: host blob in, host blob out.

    Returns (out, out_offsets, status). Output sizes come from the matching `cri_*_sizes` call (exact, header-only);
    the library pipelines the batch in chunks of whole streams (H2D / kernels / D2H overlapped).
    """
    ctx = ctx or default_context()
    L = ctx._lib
    blob = np.ascontiguousarray(blob, dtype=np.uint8)
    offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
    n = len(offsets) - 1
    sizes = np.zeros(max(n, 1), dtype=np.uint64)
    status = np.zeros(max(n, 1), dtype=np.int32)
    bp, op = blob.ctypes.data, offsets.ctypes.data
    keys = None if keys is None else np.ascontiguousarray(keys, dtype=np.uint64)
    subkeys = None if subkeys is None else np.ascontiguousarray(subkeys, dtype=np.uint16)
    kp = None if keys is None else keys.ctypes.data
    sp = None if subkeys is None else subkeys.ctypes.data
    adx = adx if adx is not None else adx_params()
    if kind == _lib.JOB_ADX_DECODE:
        L.cri_adx_decode_sizes(bp, op, n, sizes.ctypes.data, status.ctypes.data)
    elif kind == _lib.JOB_ADX_ENCODE:
        L.cri_adx_encode_sizes(bp, op, n, ctypes.byref(adx), sizes.ctypes.data, status.ctypes.data)
    elif kind == _lib.JOB_HCA_DECODE:
        L.cri_hca_decode_sizes(bp, op, n, sizes.ctypes.data, status.ctypes.data)
    elif kind == _lib.JOB_HCA_ENCODE:
        L.cri_hca_encode_sizes_ex(bp, op, n, int(quality), int(adx.force_not_looping), sizes.ctypes.data, status.ctypes.data)
    elif kind == _lib.JOB_HCA_CRYPT:
        sizes[:n] = offsets[1:] - offsets[:-1]
    else:
        raise ValueError(f"unknown job kind {kind}")
    out_offsets = np.zeros(n + 1, dtype=np.uint64)
    np.cumsum(sizes[:n], out=out_offsets[1:])
    if kind == _lib.JOB_HCA_CRYPT:
        out_offsets = offsets
    total = int(out_offsets[-1])
    if out is None:
        out = np.empty(max(total, 1), dtype=np.uint8)
    elif out.size < total:
        raise ValueError("output buffer too small")
    oo = out_offsets.ctypes.data
    if kind == _lib.JOB_ADX_DECODE:
        rc = L.cri_adx_decode_batch(ctx.handle, bp, op, n, out.ctypes.data, oo, status.ctypes.data)
    elif kind == _lib.JOB_ADX_ENCODE:
        rc = L.cri_adx_encode_batch(ctx.handle, bp, op, n, ctypes.byref(adx), out.ctypes.data, oo, status.ctypes.data)
    elif kind == _lib.JOB_HCA_DECODE:
        rc = L.cri_hca_decode_batch(ctx.handle, bp, op, n, kp, sp, out.ctypes.data, oo, status.ctypes.data)
    elif kind == _lib.JOB_HCA_ENCODE:
        rc = L.cri_hca_encode_batch(ctx.handle, bp, op, n, int(quality), int(adx.force_not_looping), out.ctypes.data, oo,
                                    status.ctypes.data)
    else:
        rc = L.cri_hca_crypt_batch(ctx.handle, bp, op, n, int(encrypt), int(ciph_type), kp, sp, out.ctypes.data, status.ctypes.data)
    ctx.check(rc)
    return out, out_offsets, status[:n]


def sizes_device(kind: int, blob, offsets: np.ndarray, ctx: Optional[Context] = None, stream=None, *, adx: Optional[AdxParams] = None,
                 quality: int = 1, **_ignored):
    """Exact output size of every stream of a device-resident blob (`cri_sizes_dev`): (sizes uint64[n], status int32[n])."""
    import torch
    ctx = ctx or default_context()
    offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
    n = len(offsets) - 1
    if stream is None:
        stream = torch.cuda.current_stream(blob.device)
    sizes = np.zeros(max(n, 1), dtype=np.uint64)
    status = np.zeros(max(n, 1), dtype=np.int32)
    adx = adx if adx is not None else adx_params()
    ctx.check(ctx._lib.cri_sizes_dev(ctx.handle, kind, blob.data_ptr(), offsets.ctypes.data, n, ctypes.byref(adx), int(quality),
                                     sizes.ctypes.data, status.ctypes.data, ctypes.c_void_p(stream.cuda_stream)))
    return sizes[:n], status[:n]


def batch_device(kind: int, blob, offsets: np.ndarray, ctx: Optional[Context] = None, out=None, stream=None, *, keys=None,
                 subkeys=None, adx: Optional[AdxParams] = None, quality: int = 1, encrypt: int = 0, ciph_type: int = 0,
                 out_offsets: Optional[np.ndarray] = None):
    """Device-resident batch: `blob` is a CUDA uint8 tensor on the context's GPU (anything with `data_ptr()`), the result
    is a CUDA uint8 tensor too -- the payload never crosses PCIe (the `cri_*_batch_dev` calls of the C-ABI).

    Returns (out, out_offsets, status). `out` may be passed in (a CUDA uint8 tensor of at least the packed output
    size); `stream` is a torch.cuda.Stream (default: the current stream of the blob's device). Headers are fetched from
    the device blob for planning (a few hundred bytes per stream). `out_offsets` (the packed layout, e.g. from
    `sizes_device`) saves the size query.
    """
    import torch
    ctx = ctx or default_context()
    L = ctx._lib
    offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
    n = len(offsets) - 1
    if stream is None:
        stream = torch.cuda.current_stream(blob.device)
    sp_ = ctypes.c_void_p(stream.cuda_stream)
    sizes = np.zeros(max(n, 1), dtype=np.uint64)
    status = np.zeros(max(n, 1), dtype=np.int32)
    keys = None if keys is None else np.ascontiguousarray(keys, dtype=np.uint64)
    subkeys = None if subkeys is None else np.ascontiguousarray(subkeys, dtype=np.uint16)
    kp = None if keys is None else keys.ctypes.data
    sp = None if subkeys is None else subkeys.ctypes.data
    adx = adx if adx is not None else adx_params()
    bp, op = blob.data_ptr(), offsets.ctypes.data
    if kind == _lib.JOB_HCA_CRYPT:
        out_offsets = offsets
    elif out_offsets is None:
        ctx.check(L.cri_sizes_dev(ctx.handle, kind, bp, op, n, ctypes.byref(adx), int(quality), sizes.ctypes.data, status.ctypes.data, sp_))
        out_offsets = np.zeros(n + 1, dtype=np.uint64)
        np.cumsum(sizes[:n], out=out_offsets[1:])
    else:
        out_offsets = np.ascontiguousarray(out_offsets, dtype=np.uint64)
    total = int(out_offsets[-1])
    if out is None:
        out = torch.empty(max(total, 1), dtype=torch.uint8, device=blob.device)
    elif out.numel() < total:
        raise ValueError("output buffer too small")
    oo = out_offsets.ctypes.data
    if kind == _lib.JOB_ADX_DECODE:
        rc = L.cri_adx_decode_batch_dev(ctx.handle, bp, op, n, out.data_ptr(), oo, status.ctypes.data, sp_)
    elif kind == _lib.JOB_ADX_ENCODE:
        rc = L.cri_adx_encode_batch_dev(ctx.handle, bp, op, n, ctypes.byref(adx), out.data_ptr(), oo, status.ctypes.data, sp_)
    elif kind == _lib.JOB_HCA_DECODE:
        rc = L.cri_hca_decode_batch_dev(ctx.handle, bp, op, n, kp, sp, out.data_ptr(), oo, status.ctypes.data, sp_)
    elif kind == _lib.JOB_HCA_ENCODE:
        rc = L.cri_hca_encode_batch_dev(ctx.handle, bp, op, n, int(quality), int(adx.force_not_looping), out.data_ptr(), oo,
                                        status.ctypes.data, sp_)
    elif kind == _lib.JOB_HCA_CRYPT:
        rc = L.cri_hca_crypt_batch_dev(ctx.handle, bp, op, n, int(encrypt), int(ciph_type), kp, sp, out.data_ptr(), status.ctypes.data, sp_)
    else:
        raise ValueError(f"unknown job kind {kind}")
    ctx.check(rc)
    return out, out_offsets, status[:n]


def _run_streams(kind: int, streams: Sequence[bytes], ctx: Optional[Context], raise_errors: bool, **kw):
    blob, offsets = pack(streams)
    out, offs, status = batch_blob(kind, blob, offsets, ctx, **kw)
    results: List[object] = []
    for i in range(len(streams)):
        if status[i] != 0:
            err = exception_for(int(status[i]))
            if raise_errors:
                raise err
            results.append(err)
        else:
            results.append(out[int(offs[i]):int(offs[i + 1])].tobytes())
    return results


def adx_params(BitDepth=4, Blocksize=0x12, Encoding=3, Highpass_Frequency=0x1F4, Filter=0, AdxVersion=4,
               force_not_looping=False) -> AdxParams:
    for name, v in (("BitDepth", BitDepth), ("Blocksize", Blocksize), ("Encoding", Encoding),
                    ("Highpass_Frequency", Highpass_Frequency), ("Filter", Filter), ("AdxVersion", AdxVersion)):
        if not 0 <= int(v) <= 0xFFFFFFFF:
            raise OverflowError(f"{name} does not fit an unsigned int")  # PyArg "I" semantics, adx.cpp:527
    return AdxParams(int(BitDepth), int(Blocksize), int(Encoding), int(Highpass_Frequency), int(Filter), int(AdxVersion),
                     1 if force_not_looping else 0)


def adx_decode_batch(streams: Sequence[bytes], ctx: Optional[Context] = None, raise_errors: bool = True):
    return _run_streams(_lib.JOB_ADX_DECODE, streams, ctx, raise_errors)


def adx_encode_batch(streams: Sequence[bytes], ctx: Optional[Context] = None, raise_errors: bool = True, **params):
    return _run_streams(_lib.JOB_ADX_ENCODE, streams, ctx, raise_errors, adx=adx_params(**params))


def _key_arrays(n: int, keys, subkeys):
    def expand(v, dtype):
        if v is None:
            return None
        if np.isscalar(v) or isinstance(v, int):
            return np.full(n, int(v), dtype=dtype)
        return np.asarray(list(v), dtype=dtype)
    return expand(keys, np.uint64), expand(subkeys, np.uint16)


def hca_decode_batch(streams: Sequence[bytes], keys=None, subkeys=None, ctx: Optional[Context] = None,
                     raise_errors: bool = True):
    k, s = _key_arrays(len(streams), keys, subkeys)
    return _run_streams(_lib.JOB_HCA_DECODE, streams, ctx, raise_errors, keys=k, subkeys=s)


def hca_crypt_batch(streams: Sequence[bytes], encrypt: bool, keys=None, subkeys=None, ciph_type: int = 56,
                    ctx: Optional[Context] = None, raise_errors: bool = True):
    k, s = _key_arrays(len(streams), keys, subkeys)
    return _run_streams(_lib.JOB_HCA_CRYPT, streams, ctx, raise_errors, keys=k, subkeys=s, encrypt=1 if encrypt else 0,
                        ciph_type=ciph_type)


def hca_encode_batch(streams: Sequence[bytes], quality: int = 1, force_not_looping: bool = False,
                     ctx: Optional[Context] = None, raise_errors: bool = True):
    p = adx_params(force_not_looping=force_not_looping)
    return _run_streams(_lib.JOB_HCA_ENCODE, streams, ctx, raise_errors, quality=int(quality), adx=p)


# host-only helpers (no GPU needed) --------------------------------------
def crc16(data: bytes) -> int:
    return int(_lib.lib().cri_crc16(bytes(data), len(data)))


def cipher_table(ciph_type: int, key: int) -> bytes:
    t = ctypes.create_string_buffer(256)
    _lib.lib().cri_hca_cipher_table(ciph_type, key, t)
    return t.raw


def mix_subkey(key: int, subkey: int) -> int:
    return int(_lib.lib().cri_hca_mix_subkey(key, subkey))


def adx_coefficients(highpass: int, rate: int):
    c = (ctypes.c_int32 * 2)()
    _lib.lib().cri_adx_coefficients(highpass, rate, c)
    return int(c[0]), int(c[1])
