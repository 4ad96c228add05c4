"""GPU: the one-call batch entry points cut a batch into chunks of whole streams and pipeline them
(copy in / kernels / copy out on three streams). Chunking must not change a byte, whatever the chunk size,
per-stream keys must follow their streams, and a caller-chosen (padded) output layout must be honoured."""
import os

import numpy as np
import pytest

from pycricodecs_b200 import _lib, engine, synth

pytestmark = pytest.mark.gpu
KEY = 0xCF222F1FE0748978


@pytest.fixture
def tiny_chunks():
    old = os.environ.get("CRI_CHUNK_MB")
    os.environ["CRI_CHUNK_MB"] = "1"
    yield
    if old is None:
        del os.environ["CRI_CHUNK_MB"]
    else:
        os.environ["CRI_CHUNK_MB"] = old


def _wavs(n):
    return [synth.wav(100 + s, 1 + s % 2, 30000 + 1234 * (s % 7)) for s in range(n)]


def test_adx_round_trip_chunked(port, ctx, tiny_chunks):
    wavs = _wavs(48)                                   # ~4.5 MB of PCM -> several chunks at 1 MB
    enc = engine.adx_encode_batch(wavs, ctx=ctx)
    for s in (0, 1, 17, 47):
        assert enc[s] == port.adx_encode(wavs[s])[1]
    dec = engine.adx_decode_batch(enc, ctx=ctx)
    for s in (0, 5, 30, 47):
        assert dec[s] == port.adx_decode(enc[s])[1]


def test_hca_chunked_keys_follow_streams(port, ctx, tiny_chunks):
    wavs = _wavs(40)
    plain = engine.hca_encode_batch(wavs, quality=1, ctx=ctx)
    for s in (0, 13, 39):
        assert plain[s] == port.hca_encode(wavs[s], 1)[1]
    keys = [KEY + 977 * s for s in range(len(plain))]
    subs = [(s * 4099) & 0xFFFF for s in range(len(plain))]
    enc = engine.hca_crypt_batch(plain, True, keys=keys, subkeys=subs, ctx=ctx)
    for s in (0, 7, 21, 39):
        assert enc[s] == port.hca_crypt(plain[s], 1, 56, keys[s], subs[s])[1]
    pcm = engine.hca_decode_batch(enc, keys=keys, subkeys=subs, ctx=ctx)
    want = engine.hca_decode_batch(plain, ctx=ctx)
    assert pcm == want
    for s in (0, 22, 39):
        assert pcm[s] == port.hca_decode(plain[s])[1]


def test_chunked_equals_unchunked(ctx):
    wavs = _wavs(36)
    whole = engine.adx_encode_batch(wavs, ctx=ctx)
    os.environ["CRI_CHUNK_MB"] = "1"
    try:
        cut = engine.adx_encode_batch(wavs, ctx=ctx)
    finally:
        del os.environ["CRI_CHUNK_MB"]
    assert cut == whole


def test_padded_output_layout_and_bad_stream(port, ctx, tiny_chunks):
    """Caller-chosen layout with gaps (not the packed one) + a corrupt stream in the middle of a chunked batch."""
    wavs = _wavs(30)
    wavs[11] = b"RIFX" + wavs[11][4:]
    blob, offsets = engine.pack(wavs)
    L = ctx._lib
    n = len(wavs)
    sizes = np.zeros(n, np.uint64)
    status = np.zeros(n, np.int32)
    p = engine.adx_params()
    import ctypes
    L.cri_adx_encode_sizes(blob.ctypes.data, offsets.ctypes.data, n, ctypes.byref(p), sizes.ctypes.data, status.ctypes.data)
    assert status[11] != 0 and sizes[11] == 0
    out_off = np.zeros(n + 1, np.uint64)
    np.cumsum(sizes + np.uint64(64), out=out_off[1:])          # 64 spare bytes after every stream
    out = np.full(int(out_off[-1]), 0xEE, np.uint8)
    st2 = np.zeros(n, np.int32)
    ctx.check(L.cri_adx_encode_batch(ctx.handle, blob.ctypes.data, offsets.ctypes.data, n, ctypes.byref(p), out.ctypes.data,
                                     out_off.ctypes.data, st2.ctypes.data))
    assert st2[11] == status[11] and int((st2 != 0).sum()) == 1
    for s in (0, 10, 12, 29):
        a, sz = int(out_off[s]), int(sizes[s])
        assert out[a:a + sz].tobytes() == port.adx_encode(wavs[s])[1]
        assert (out[a + sz:a + sz + 64] == 0xEE).all()


def test_pool_reuse_and_trim(ctx):
    wavs = _wavs(8)
    a = engine.adx_encode_batch(wavs, ctx=ctx)
    b = engine.adx_encode_batch(wavs, ctx=ctx)          # second call runs out of the context's HBM cache
    assert a == b
    ctx._lib.cri_ctx_trim(ctx.handle)
    assert engine.adx_encode_batch(wavs, ctx=ctx) == a


def test_pinned_host_buffers(ctx):
    L = ctx._lib
    p = L.cri_host_alloc(1 << 20)
    assert p
    L.cri_host_free(p)
    assert _lib.JOB_ADX_DECODE is not None
