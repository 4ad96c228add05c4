"""WAV ingest breadth: 8 / 20 / 24 / 32-bit integer, float32 / float64 and WAVE_FORMAT_EXTENSIBLE
images go through the reference's PCM16 conversion (PCM::load_WAVE / Get_PCM16, pcm.cpp:291-327, 455-545) -- here a
device kernel in front of the encoders. Expected values: the compiled reference's own encodes of the same images
(tools/make_golden_ingest.py -> tests/golden/ingest_digests.json)."""
import ctypes
import hashlib
import json
import os
import struct

import numpy as np
import pytest

from helpers import wavgen

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ingest_digests.json")
h = lambda b: hashlib.sha256(b).hexdigest()[:16]


def _cases():
    d = json.load(open(GOLD))
    wavs = [wavgen.wav_as(k, sid, ch, n) for k, sid, ch, n in wavgen.CASES]
    for (k, *_), w in zip(wavgen.CASES, wavs):
        assert h(w) == d[k]["wav"], f"generator drifted for {k}"
    return d, wavs


def test_output_sizes_follow_the_sample_encoding():
    """Header-only size queries (no GPU): sample counts come from data bytes / container size."""
    from pycricodecs_b200 import _lib, engine
    d, wavs = _cases()
    L = _lib.lib()
    blob, off = engine.pack(wavs)
    n = len(wavs)
    sizes, status = np.zeros(n, np.uint64), np.zeros(n, np.int32)
    p = engine.adx_params()
    assert L.cri_adx_encode_sizes(blob.ctypes.data, off.ctypes.data, n, ctypes.byref(p), sizes.ctypes.data, status.ctypes.data) == 0
    assert status.tolist() == [0] * n and sizes.tolist() == [d[k]["adx_len"] for k, *_ in wavgen.CASES]
    assert L.cri_hca_encode_sizes(blob.ctypes.data, off.ctypes.data, n, 1, sizes.ctypes.data, status.ctypes.data) == 0
    assert status.tolist() == [0] * n and sizes.tolist() == [d[k]["hca_len"] for k, *_ in wavgen.CASES]
    # combinations the reference's loader cannot convert are refused with its error -8 (pcm.cpp:32)
    bad = wavgen._riff(struct.pack("<HHIIHH", 1, 1, 48000, 96000, 2, 8), bytes(64))       # 8 valid bits in a 2-byte container
    blob, off = engine.pack([bad])
    assert L.cri_adx_encode_sizes(blob.ctypes.data, off.ctypes.data, 1, ctypes.byref(p), sizes.ctypes.data, status.ctypes.data) == 0
    assert status[0] == -108


@pytest.mark.gpu
def test_every_sample_encoding_encodes_like_the_reference(ctx):
    from pycricodecs_b200 import engine
    d, wavs = _cases()
    adx = engine.adx_encode_batch(wavs, ctx=ctx)
    hca = engine.hca_encode_batch(wavs, quality=1, ctx=ctx)
    for (k, *_), a, x in zip(wavgen.CASES, adx, hca):
        assert (h(a), len(a)) == (d[k]["adx"], d[k]["adx_len"]), f"ADX, {k}"
        assert (h(x), len(x)) == (d[k]["hca"], d[k]["hca_len"]), f"HCA, {k}"
    # mixed with plain 16-bit streams in one batch
    from pycricodecs_b200 import synth
    plain = synth.wav(3, 2, 2048)
    got = engine.adx_encode_batch([plain, wavs[1], plain, wavs[4]], ctx=ctx)
    assert got[0] == got[2] and h(got[1]) == d["s24"]["adx"] and h(got[3]) == d["f32"]["adx"]


@pytest.mark.gpu
def test_looping_wav_to_adx_and_back(ctx):
    """A WAV with a sampler loop encodes an ADX loop table (Loop::writeLoops, adx.cpp:94-143); decoding that ADX gives
    a WAV with the loop in a smpl chunk again. Expected values: the compiled reference (adx_loops in the fixture)."""
    from pycricodecs_b200 import engine
    d = json.load(open(GOLD))["adx_loops"]
    for (sid, ch, n, ls, le, ver), want in zip(wavgen.LOOP_CASES, d):
        w = wavgen.loop_wav(sid, ch, n, ls, le)
        assert h(w) == want["wav"]
        a = engine.adx_encode_batch([w], ctx=ctx, AdxVersion=ver)[0]
        assert (h(a), len(a)) == (want["adx"], want["adx_len"]), f"encode, case {sid}"
        if not want["adx_decoded"].startswith("error"):
            assert h(engine.adx_decode_batch([a], ctx=ctx)[0]) == want["adx_decoded"], f"decode, case {sid}"


@pytest.mark.gpu
def test_looping_wav_to_hca_and_back(ctx):
    """Looping HCA encode: loop chunk in the header, the loop start frame moved to a 2048-byte boundary with extra
    delay frames, and the reference's frame feeder (silence / first-sample pre-roll, main audio up to the loop end,
    post-roll taken from the loop start; hca.cpp:2292-2321, 2440-2449, 3000-3107) -- assembled on the device in front of
    the encode kernel. Decoding that stream puts the loop back into a smpl chunk."""
    from pycricodecs_b200 import engine
    d = json.load(open(GOLD))["hca_loops"]
    wavs = [wavgen.loop_wav(sid, ch, n, ls, le) for sid, ch, n, ls, le, q in wavgen.HCA_LOOP_CASES]
    for (sid, ch, n, ls, le, q), w, want in zip(wavgen.HCA_LOOP_CASES, wavs, d):
        assert h(w) == want["wav"]
        x = engine.hca_encode_batch([w], quality=q, ctx=ctx)[0]
        assert (h(x), len(x)) == (want["hca"], want["hca_len"]), f"encode, case {sid}"
        assert h(engine.hca_decode_batch([x], ctx=ctx)[0]) == want["hca_decoded"], f"decode, case {sid}"
    # force_not_looping ignores the smpl chunk: same stream as the loop-free WAV
    from pycricodecs_b200 import synth
    sid, ch, n, ls, le, q = wavgen.HCA_LOOP_CASES[0]
    assert engine.hca_encode_batch([wavs[0]], quality=q, force_not_looping=True, ctx=ctx)[0] == \
        engine.hca_encode_batch([synth.wav(sid, ch, n)], quality=q, ctx=ctx)[0]
