"""CPU: the C-ABI library loads, exports every symbol include/cricodecs_b200.h declares, and its host-only
entry points (no device needed) agree with the oracle. Compute entry points must FAIL without a GPU."""
import ctypes
import os
import re

import numpy as np
import pytest

from pycricodecs_b200 import _lib, engine, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEY = 0xCF222F1FE0748978


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "cricodecs_b200.h")).read()
    return sorted(set(re.findall(r"CRI_API [^;(]*?\b(cri_\w+)\s*\(", text)))


def test_every_declared_symbol_is_exported_and_bound():
    names = declared_symbols()
    assert len(names) >= 30
    L = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(L, n), f"{n} declared in the header but not exported"
    assert set(names) == set(_lib.SIGNATURES), "python binding table and header disagree"


def test_host_helpers_match_oracle(port):
    assert engine.crc16(b"\xff\xff\x12\x34") == port.crc16(b"\xff\xff\x12\x34") == 0xEC9F
    rng = np.random.default_rng(0)
    for n in (1, 2, 17, 682, 1024):
        b = rng.integers(0, 256, n, dtype=np.uint8).tobytes()
        assert engine.crc16(b) == port.crc16(b)
    for t, k in ((56, KEY), (56, 1), (56, 0), (1, 0), (0, 5), (56, 0xFFFFFFFFFFFFFFFF)):
        assert engine.cipher_table(t, k) == port.cipher_table(t, k)
    for k, s in ((KEY, 0), (KEY, 0x1234), (1, 0xFFFF), (0, 7)):
        assert engine.mix_subkey(k, s) == port.mix_subkey(k, s)
    for hp, rate in ((500, 48000), (500, 44100), (0, 48000), (65535, 8000), (100, 22050)):
        assert engine.adx_coefficients(hp, rate) == port.adx_coefficients(hp, rate)


def _sizes(fn, streams, *extra):
    blob, off = engine.pack(streams)
    sizes = np.zeros(len(streams), np.uint64)
    status = np.zeros(len(streams), np.int32)
    rc = fn(blob.ctypes.data, off.ctypes.data, len(streams), *extra, sizes.ctypes.data, status.ctypes.data)
    assert rc == 0
    return sizes.tolist(), status.tolist()


def test_size_queries_match_oracle_outputs(port):
    L = _lib.lib()
    wavs = [synth.wav(0, 2, 3200), synth.wav(1, 1, 1000), b"RIFFxxxxWAVE", synth.wav(2, 2, 31)]
    p = engine.adx_params()
    sizes, status = _sizes(L.cri_adx_encode_sizes, wavs, ctypes.byref(p))
    for w, s, st in zip(wavs, sizes, status):
        r, a = port.adx_encode(w)
        assert (st == 0) == (r == 0)
        if r == 0:
            assert s == len(a)
    for q in range(4):
        sizes, status = _sizes(L.cri_hca_encode_sizes, wavs, q)
        for w, s, st in zip(wavs, sizes, status):
            r, x = port.hca_encode(w, q)
            assert (st == 0) == (r == 0)
            if r == 0:
                assert s == len(x)
    adx = [port.adx_encode(w)[1] for w in (wavs[0], wavs[1])] + [b"\x80\x00" + b"\0" * 40]
    sizes, status = _sizes(L.cri_adx_decode_sizes, adx)
    assert status[2] != 0 and sizes[:2] == [len(port.adx_decode(a)[1]) for a in adx[:2]]
    hca = [port.hca_encode(wavs[0], 1)[1], port.hca_encode(wavs[1], 3)[1], b"HCA\0" + b"\0" * 100]
    sizes, status = _sizes(L.cri_hca_decode_sizes, hca)
    assert status[2] == -201 and sizes[:2] == [len(port.hca_decode(x)[1]) for x in hca[:2]]


def test_error_strings_are_the_reference_messages():
    assert engine.strerror(-12) == "Blocksize must be between 3 and 255 inclusive."
    assert engine.strerror(-202) == "Decoding error, either an incorrect key or an unknown exception."
    assert engine.strerror(-101) == "Invalid WAVE file header."
    assert isinstance(engine.exception_for(-3), NotImplementedError)
    assert isinstance(engine.exception_for(-9), ValueError)


def test_compute_fails_loudly_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        engine.Context(0)
    h = ctypes.c_void_p()
    assert _lib.lib().cri_ctx_create(0, ctypes.byref(h)) == -400


def test_hca_header_sniffing_matches_reference_fields(port):
    from pycricodecs_b200 import HCA, CriHcaQuality
    x = port.hca_encode(synth.wav(0, 2, 5000), 3)[1]
    info = HCA(x).info()
    assert info["FrameCount"] == 6 and info["FrameSize"] == 341 and info["ChannelCount"] == 2 and info["SampleRate"] == 48000
    assert info["TotalBandCount"] == 128 and info["BaseBandCount"] == 43 and info["StereoBandCount"] == 42 and info["BandsPerHfrGroup"] == 6
    e = port.hca_crypt(x, 1, 56, KEY)[1]
    o = HCA(e)
    assert o.encrypted and o.key == KEY          # default key for an encrypted stream without a key (hca.py:91-92)
    w = HCA(synth.wav(0, 2, 100))
    assert w.filetype == "wav" and w.info()["fmtChannelCount"] == 2
    with pytest.raises(ValueError, match="Invalid HCA or WAV"):
        HCA(b"nothing useful here....")
    assert [q.value for q in CriHcaQuality] == [0, 1, 2, 3, 5]


def test_size_queries_reject_what_the_planner_rejects():
    """Headers that lie (advisor findings, round 1): band counts that make the reference's HFR group count wrap, delay +
    padding beyond the stream, frames that are not there, an ADX sample count the payload cannot back."""
    import ctypes
    import numpy as np
    import sys, os
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__))))
    from helpers import hca3gen
    from pycricodecs_b200 import _lib, engine, synth
    L = _lib.lib()

    def hca_sizes(stream):
        blob, off = engine.pack([stream])
        sizes, st = np.zeros(1, np.uint64), np.zeros(1, np.int32)
        L.cri_hca_decode_sizes(blob.ctypes.data, off.ctypes.data, 1, sizes.ctypes.data, st.ctypes.data)
        return int(sizes[0]), int(st[0])

    good = hca3gen.stream(seed=1, version=0x0200, min_res=1, frames=3, frame_size=2048)
    assert hca_sizes(good) == (44 + (3 * 1024 - 128) * 4, 0)
    wrap = hca3gen.header(version=0x0200, min_res=1, frames=3, total=20, base=30, stereo=10, bands_per_hfr=1) + bytes(3 * 1024)
    assert hca_sizes(wrap) == (0, -201)
    lying = hca3gen.header(version=0x0200, min_res=1, frames=1, delay=900, padding=900) + bytes(1024)
    assert hca_sizes(lying) == (0, -201)
    assert hca_sizes(good[:-100]) == (0, -201)                       # last frame cut short

    wav = synth.wav(2, 1, 320)
    # a hand-made ADX header (adx.cpp:145-183): 10 blocks of payload, sample count 2^31
    hdr = bytearray(b"\x80\x00" + (0x20 - 4).to_bytes(2, "big") + bytes([3, 18, 4, 1]) + (48000).to_bytes(4, "big") + (1 << 31).to_bytes(4, "big")
                    + (500).to_bytes(2, "big") + bytes([3, 0]))
    hdr += bytes(0x20 - 6 - len(hdr)) + b"(c)CRI"
    adx = bytes(hdr) + bytes(18 * 10)
    blob, off = engine.pack([adx])
    sizes, st = np.zeros(1, np.uint64), np.zeros(1, np.int32)
    L.cri_adx_decode_sizes(blob.ctypes.data, off.ctypes.data, 1, sizes.ctypes.data, st.ctypes.data)
    assert (int(sizes[0]), int(st[0])) == (0, -301)
    assert wav[:4] == b"RIFF"


def test_size_queries_on_host_threads_equal_the_serial_answers(port):
    """Batches of 512 streams and more are parsed by several host threads (formats.h: parallel_for). The size queries run
    without a GPU: a mixed batch of 700 streams (good ones of different lengths, truncated ones, garbage) must give, per
    stream, what a one-stream query gives."""
    import numpy as np
    from pycricodecs_b200 import _lib, engine, synth
    L = _lib.lib()
    wavs = [synth.wav(40 + k, 1 + k % 2, 1500 + 333 * k) for k in range(5)]
    hcas = [port.hca_encode(w, 1 + k % 3)[1] for k, w in enumerate(wavs)]
    pool = hcas + [hcas[0][:-50], b"not an hca stream at all", hcas[3][:40]]
    streams = [pool[(7 * i + i // 5) % len(pool)] for i in range(700)]
    blob, off = engine.pack(streams)
    sizes, st = np.zeros(len(streams), np.uint64), np.zeros(len(streams), np.int32)
    assert L.cri_hca_decode_sizes(blob.ctypes.data, off.ctypes.data, len(streams), sizes.ctypes.data, st.ctypes.data) == 0
    single = {}
    for k, s in enumerate(pool):
        b1, o1 = engine.pack([s])
        z, t = np.zeros(1, np.uint64), np.zeros(1, np.int32)
        L.cri_hca_decode_sizes(b1.ctypes.data, o1.ctypes.data, 1, z.ctypes.data, t.ctypes.data)
        single[k] = (int(z[0]), int(t[0]))
    assert len({v for v in single.values()}) >= 5
    for i in range(len(streams)):
        assert (int(sizes[i]), int(st[i])) == single[(7 * i + i // 5) % len(pool)], i
    # WAV -> HCA sizes take the same threaded route
    wpool = wavs + [wavs[1][:30], b"RIFFxxxxWAVEjunk"]
    wstreams = [wpool[(3 * i) % len(wpool)] for i in range(600)]
    blob, off = engine.pack(wstreams)
    sizes, st = np.zeros(len(wstreams), np.uint64), np.zeros(len(wstreams), np.int32)
    assert L.cri_hca_encode_sizes(blob.ctypes.data, off.ctypes.data, len(wstreams), 1, sizes.ctypes.data, st.ctypes.data) == 0
    for k, w in enumerate(wpool):
        b1, o1 = engine.pack([w])
        z, t = np.zeros(1, np.uint64), np.zeros(1, np.int32)
        L.cri_hca_encode_sizes(b1.ctypes.data, o1.ctypes.data, 1, 1, z.ctypes.data, t.ctypes.data)
        for i in range(k, len(wstreams), len(wpool)):
            if (3 * i) % len(wpool) == k:
                assert (int(sizes[i]), int(st[i])) == (int(z[0]), int(t[0])), (i, k)
