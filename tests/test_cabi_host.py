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
