"""GPU, BASELINE.json sizes (8192 x 2 s 48 kHz stereo streams per GPU): sampled bit-exactness against the compiled reference (oracle/_ref, where present; else its C restatement) plus
size-independent properties over the WHOLE batch -- results must not depend on how the batch is tiled, an
encrypt -> decrypt round trip must return every byte, decrypt+decode must equal plain decode, and a batch must equal
the concatenation of its halves."""
import hashlib
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

pytestmark = pytest.mark.gpu
KEY = 0xCF222F1FE0748978
STREAMS = 8192
SAMPLE = (0, 1, 2, 511, 512, 1000, 4095, 4096, 5000, 8190, 8191) + tuple(range(37, 8192, 409))


def _digest(blob, offsets):
    """checksum of per-stream checksums"""
    h = hashlib.sha256()
    for i in range(len(offsets) - 1):
        h.update(hashlib.sha256(blob[int(offsets[i]):int(offsets[i + 1])]).digest())
    return h.hexdigest()


def _run(ctx, kind, blob, offsets, **kw):
    from pycricodecs_b200 import engine
    with engine.Job(ctx, kind, blob, offsets, **kw) as job:
        job.run()
        out = np.empty(max(job.out_bytes, 1), np.uint8)
        _, status = job.download(out)
        assert int((status != 0).sum()) == 0
        return out[: job.out_bytes], job.out_offsets.copy()


@pytest.fixture(scope="module")
def corpus(ctx):
    import torch
    import bench
    wav, woff = bench.make_wav_blob(STREAMS, 0, torch.device("cuda", 0))
    return wav.numpy()[: int(woff[-1])], woff


def _stream(blob, off, i):
    return bytes(blob[int(off[i]):int(off[i + 1])])


def test_hca_full_batch(checker, ctx, corpus, monkeypatch):
    from pycricodecs_b200 import _lib, engine
    wav, woff = corpus
    hca, hoff = _run(ctx, _lib.JOB_HCA_ENCODE, wav, woff, quality=1, adx=engine.adx_params())
    assert len(hoff) == STREAMS + 1 and int(hoff[-1]) == STREAMS * 64204
    for i in SAMPLE:
        assert _stream(hca, hoff, i) == checker.hca_encode(_stream(wav, woff, i), 1), f"encode, stream {i} ({checker.kind})"
    pcm, poff = _run(ctx, _lib.JOB_HCA_DECODE, hca, hoff, keys=None)
    assert int(poff[-1]) == STREAMS * 384044
    for i in SAMPLE:
        assert _stream(pcm, poff, i) == checker.hca_decode(_stream(hca, hoff, i)), f"decode, stream {i} ({checker.kind})"
    whole = _digest(pcm, poff)
    # tiling independence: another run length (runs cross stream boundaries elsewhere), and the general kernels
    monkeypatch.setenv("CRI_HCA_FAST_RUN", "7")
    pcm2, poff2 = _run(ctx, _lib.JOB_HCA_DECODE, hca, hoff, keys=None)
    assert np.array_equal(poff, poff2) and _digest(pcm2, poff2) == whole
    monkeypatch.delenv("CRI_HCA_FAST_RUN")
    half = STREAMS // 2
    monkeypatch.setenv("CRI_HCA_GENERAL", "1")
    pcm3, poff3 = _run(ctx, _lib.JOB_HCA_DECODE, hca[: int(hoff[half])], hoff[: half + 1], keys=None)
    monkeypatch.delenv("CRI_HCA_GENERAL")
    assert np.array_equal(pcm3, pcm[: int(poff[half])])
    # encrypt -> decrypt returns every byte; decrypt+decode equals plain decode
    keys = np.full(STREAMS, KEY, np.uint64)
    enc, eoff = _run(ctx, _lib.JOB_HCA_CRYPT, hca, hoff, keys=keys, encrypt=1, ciph_type=56)
    assert np.array_equal(eoff, hoff) and not np.array_equal(enc, hca)
    for i in SAMPLE[:6]:
        assert _stream(enc, eoff, i) == checker.hca_crypt(_stream(hca, hoff, i), 1, 56, KEY)
    dec, doff = _run(ctx, _lib.JOB_HCA_CRYPT, enc, eoff, keys=keys, encrypt=0, ciph_type=0)
    assert np.array_equal(dec, hca)
    pcm4, poff4 = _run(ctx, _lib.JOB_HCA_DECODE, enc, eoff, keys=keys)
    assert _digest(pcm4, poff4) == whole


def test_adx_full_batch(checker, ctx, corpus):
    from pycricodecs_b200 import _lib, engine
    wav, woff = corpus
    adx, aoff = _run(ctx, _lib.JOB_ADX_ENCODE, wav, woff, adx=engine.adx_params())
    assert int(aoff[-1]) == STREAMS * 108066
    for i in SAMPLE:
        assert _stream(adx, aoff, i) == checker.adx_encode(_stream(wav, woff, i)), f"encode, stream {i} ({checker.kind})"
    pcm, poff = _run(ctx, _lib.JOB_ADX_DECODE, adx, aoff)
    assert int(poff[-1]) == STREAMS * 384044
    for i in SAMPLE:
        assert _stream(pcm, poff, i) == checker.adx_decode(_stream(adx, aoff, i)), f"decode, stream {i} ({checker.kind})"
    # a batch equals the concatenation of its halves (chains are independent; tiles / groups / CTAs differ)
    half = STREAMS // 2
    a2, o2 = _run(ctx, _lib.JOB_ADX_ENCODE, wav[int(woff[half]):], woff[half:] - woff[half], adx=engine.adx_params())
    assert np.array_equal(a2, adx[int(aoff[half]):])
    p2, _ = _run(ctx, _lib.JOB_ADX_DECODE, adx[: int(aoff[half])], aoff[: half + 1])
    assert np.array_equal(p2, pcm[: int(poff[half])])
