"""GPU: the device-pointer batch calls (`cri_*_batch_dev`): input and output blobs live in
HBM, work is ordered on the caller's stream, planning reads only the headers it fetches from the device blob. Results must
be the bytes of the host-buffer calls (and of the oracle); WAVs whose chunks sit in unusual places must still parse."""
import struct

import numpy as np
import pytest

from pycricodecs_b200 import _lib, engine, synth

pytestmark = pytest.mark.gpu
KEY = 0xCF222F1FE0748978


def _torch():
    import torch
    return torch


def _to_dev(streams):
    torch = _torch()
    blob, offsets = engine.pack(streams)
    return torch.from_numpy(blob.copy()).cuda(), offsets


def _split(out, offs):
    host = out.cpu().numpy()
    return [host[int(offs[i]):int(offs[i + 1])].tobytes() for i in range(len(offs) - 1)]


def _wavs(n):
    return [synth.wav(300 + s, 1 + s % 2, 20000 + 777 * (s % 5)) for s in range(n)]


def test_hca_round_trip_on_device(port, ctx):
    torch = _torch()
    wavs = _wavs(12)
    d_wav, woff = _to_dev(wavs)
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        hca_t, hoff, st = engine.batch_device(_lib.JOB_HCA_ENCODE, d_wav, woff, ctx, quality=1)
    assert not st.any()
    hcas = _split(hca_t, hoff)
    for s in (0, 5, 11):
        assert hcas[s] == port.hca_encode(wavs[s], 1)[1]
    keys = np.full(len(wavs), KEY, np.uint64)
    enc_t, eoff, st = engine.batch_device(_lib.JOB_HCA_CRYPT, hca_t, hoff, ctx, keys=keys, encrypt=1, ciph_type=56)
    assert not st.any()
    assert _split(enc_t, eoff) == engine.hca_crypt_batch(hcas, True, keys=KEY, ctx=ctx)
    pcm_t, poff, st = engine.batch_device(_lib.JOB_HCA_DECODE, enc_t, eoff, ctx, keys=keys)
    assert not st.any()
    pcm = _split(pcm_t, poff)
    assert pcm == engine.hca_decode_batch(hcas, ctx=ctx)
    assert pcm[3] == port.hca_decode(hcas[3])[1]


def test_adx_round_trip_on_device(port, ctx):
    wavs = _wavs(10)
    d_wav, woff = _to_dev(wavs)
    adx_t, aoff, st = engine.batch_device(_lib.JOB_ADX_ENCODE, d_wav, woff, ctx)
    assert not st.any()
    adx = _split(adx_t, aoff)
    assert adx == engine.adx_encode_batch(wavs, ctx=ctx)
    assert adx[7] == port.adx_encode(wavs[7])[1]
    pcm_t, poff, st = engine.batch_device(_lib.JOB_ADX_DECODE, adx_t, aoff, ctx)
    assert _split(pcm_t, poff) == engine.adx_decode_batch(adx, ctx=ctx)


def _wav_with_chunks(sid, front_junk, tail_loop):
    """A WAV whose `data` chunk sits behind `front_junk` bytes of an unknown chunk and, optionally, in front of a sampler loop."""
    base = synth.wav(sid, 2, 6000)
    fmt, data = base[12:36], base[36:]
    junk = b"JUNK" + struct.pack("<I", front_junk) + bytes(front_junk) if front_junk else b""
    smpl = b""
    if tail_loop:
        body = struct.pack("<9I", 0, 0, 0, 60, 0, 0, 0, 1, 0) + struct.pack("<6I", 0, 0, 1024, 5000, 0, 0)
        smpl = b"smpl" + struct.pack("<I", len(body)) + body
    payload = b"WAVE" + fmt + junk + data + smpl
    return b"RIFF" + struct.pack("<I", len(payload)) + payload


def test_wav_headers_beyond_the_first_fetch(ctx):
    """Chunks behind the first kilobyte (a long chunk in front of the samples) and behind the samples (a sampler loop)."""
    wavs = [_wav_with_chunks(1, 0, False), _wav_with_chunks(2, 3000, False), _wav_with_chunks(3, 0, True),
            _wav_with_chunks(4, 5000, True), _wav_with_chunks(5, 700, True)]
    d_wav, woff = _to_dev(wavs)
    for kind, host in ((_lib.JOB_ADX_ENCODE, engine.adx_encode_batch), (_lib.JOB_HCA_ENCODE, engine.hca_encode_batch)):
        out_t, ooff, st = engine.batch_device(kind, d_wav, woff, ctx)
        want = host(wavs, ctx=ctx, raise_errors=False)
        got = _split(out_t, ooff)
        assert int((st == 0).sum()) >= 3                     # the loop-free ones at least
        for s in range(len(wavs)):
            if isinstance(want[s], Exception):
                assert st[s] == want[s].status
            else:
                assert st[s] == 0 and got[s] == want[s]


def test_bad_streams_and_status_on_device(ctx):
    wavs = _wavs(6)
    hcas = engine.hca_encode_batch(wavs, ctx=ctx)
    broken = bytearray(hcas[2])
    broken[200] ^= 0x55                                         # frame CRC fails -> -202 on the device, silence out
    hcas[2] = bytes(broken)
    hcas[4] = b"XXXX" + hcas[4][4:]                             # header fails on the host -> -201, no output bytes
    d_hca, hoff = _to_dev(hcas)
    out_t, ooff, st = engine.batch_device(_lib.JOB_HCA_DECODE, d_hca, hoff, ctx)
    want = engine.hca_decode_batch(hcas, ctx=ctx, raise_errors=False)
    assert [int(x) for x in st] == [0, 0, -202, 0, -201, 0]
    got = _split(out_t, ooff)
    assert ooff[5] == ooff[4]
    assert not any(got[2][44:])
    for s in (0, 1, 3, 5):
        assert got[s] == want[s]


def test_wrong_layout_is_refused(ctx):
    import ctypes
    torch = _torch()
    wavs = _wavs(3)
    d_wav, woff = _to_dev(wavs)
    out = torch.empty(1 << 20, dtype=torch.uint8, device="cuda")
    bad = np.array([0, 10, 20, 30], np.uint64)
    st = np.zeros(3, np.int32)
    p = engine.adx_params()
    rc = ctx._lib.cri_adx_encode_batch_dev(ctx.handle, d_wav.data_ptr(), woff.ctypes.data, 3, ctypes.byref(p), out.data_ptr(),
                                           bad.ctypes.data, st.ctypes.data, None)
    assert rc == -301


@pytest.mark.parametrize("pieces", [1, 3, 5])
def test_device_call_in_pieces_equals_the_host_call(ctx, monkeypatch, pieces):
    """Large HCA batches run as several pieces inside one `_dev` call (header fetch / planning of a piece beside the
    kernels of the piece in front). CRI_DEV_PIECES forces that on a small batch: bytes and per-stream status must not
    depend on the cut, ragged stream sizes and a corrupt stream in a later piece included (it is silenced, its status
    lands on its own index)."""
    torch = _torch()
    wavs = [synth.wav(900 + s, 2, 3000 + 1234 * (s % 7)) for s in range(23)]           # all stereo: the fast decode kernels
    hcas = engine.hca_encode_batch(wavs, quality=1, ctx=ctx)
    bad = bytearray(hcas[17])
    bad[len(bad) // 2] ^= 0x5A                                   # breaks one frame's CRC
    hcas[17] = bytes(bad)
    d_hca, hoff = _to_dev(hcas)
    monkeypatch.setenv("CRI_DEV_PIECES", str(pieces))
    pcm_t, poff, st = engine.batch_device(_lib.JOB_HCA_DECODE, d_hca, hoff, ctx)
    got = _split(pcm_t, poff)
    monkeypatch.setenv("CRI_DEV_PIECES", "1")
    ref_t, roff, rst = engine.batch_device(_lib.JOB_HCA_DECODE, d_hca, hoff, ctx)
    assert np.array_equal(poff, roff) and np.array_equal(st, rst)
    assert got == _split(ref_t, roff)
    assert st[17] != 0 and not np.delete(st, 17).any()
    assert set(got[17][44:]) <= {0}                               # a failed stream leaves silence behind its header
    good = [h for i, h in enumerate(hcas) if i != 17]
    assert [g for i, g in enumerate(got) if i != 17] == engine.hca_decode_batch(good, ctx=ctx)
    # the encoder takes the same route
    d_wav, woff = _to_dev(wavs)
    monkeypatch.setenv("CRI_DEV_PIECES", str(pieces))
    hca_t, eoff, est = engine.batch_device(_lib.JOB_HCA_ENCODE, d_wav, woff, ctx, quality=1)
    assert not est.any()
    clean = engine.hca_encode_batch(wavs, quality=1, ctx=ctx)
    assert _split(hca_t, eoff) == clean
