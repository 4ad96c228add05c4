"""GPU parity: HCA v2.0 encode kernel (through the C-ABI) vs the oracle, byte-exact bitstreams."""
import numpy as np
import pytest

from pycricodecs_b200 import HCA, CriHcaQuality, engine, synth

pytestmark = pytest.mark.gpu
Q = {0: CriHcaQuality.Highest, 1: CriHcaQuality.High, 2: CriHcaQuality.Middle, 3: CriHcaQuality.Low}


def _frame_diff(a: bytes, b: bytes):
    if len(a) != len(b):
        return f"len {len(a)} vs {len(b)}"
    if a == b:
        return None
    hs = int.from_bytes(b[6:8], "big"); fs = int.from_bytes(b[28:30], "big")
    bad = [i for i in range((len(b) - hs) // fs) if a[hs + i * fs: hs + (i + 1) * fs] != b[hs + i * fs: hs + (i + 1) * fs]]
    return f"header {'ok' if a[:hs] == b[:hs] else 'DIFF'}, {len(bad)} frames differ, first {bad[:5]}"


@pytest.mark.parametrize("quality", [0, 1, 2, 3])
@pytest.mark.parametrize("channels", [1, 2])
def test_encode_matches_oracle(port, ctx, quality, channels):
    wavs = [synth.wav(s, channels, 40000 + 1000 * s) for s in range(3)]
    got = HCA.encode_batch(wavs, Q[quality], ctx=ctx)
    for w, g in zip(wavs, got):
        r, want = port.hca_encode(w, quality)
        assert r == 0
        assert _frame_diff(g, want) is None


def test_full_length_streams(port, ctx):
    wavs = [synth.wav(s, 2) for s in (0, 1)]
    got = HCA.encode_batch(wavs, ctx=ctx)
    for w, g in zip(wavs, got):
        assert _frame_diff(g, port.hca_encode(w, 1)[1]) is None


def test_edge_lengths_and_silence(port, ctx):
    cases = [synth.wav(0, 2, 1), synth.wav(1, 1, 127), synth.wav(2, 2, 896), synth.wav(3, 2, 897), synth.wav(4, 1, 1024 * 3)]
    silent = synth.wav_header(2, 5000) + bytes(5000 * 4)
    loud = synth.wav_header(2, 3000) + (np.tile(np.array([32767, -32768], np.int16), 3000)).tobytes()
    cases += [silent, loud]
    for q in (1, 3):
        got = HCA.encode_batch(cases, Q[q], ctx=ctx)
        for w, g in zip(cases, got):
            assert _frame_diff(g, port.hca_encode(w, q)[1]) is None


def test_many_channels(port, ctx):
    n = 6000
    for ch in (3, 4, 6, 8):
        pcm = np.stack([synth.channel_samples(11, c, n) for c in range(ch)], axis=1)
        w = synth.wav_header(ch, n) + pcm.tobytes()
        for q in (1, 3):
            r, want = port.hca_encode(w, q)
            got = HCA.encode_batch([w], Q[q], ctx=ctx, raise_errors=False)[0]
            if r == 0:
                assert _frame_diff(got, want) is None, (ch, q)
            else:
                assert isinstance(got, Exception)


def test_encode_decode_round_trip_on_gpu(port, ctx):
    w = synth.wav(6, 2, 30000)
    h = HCA.encode_batch([w], ctx=ctx)[0]
    d = HCA.decode_batch([h], ctx=ctx)[0]
    assert d == port.hca_decode(port.hca_encode(w, 1)[1])[1]
    pcm_in = np.frombuffer(w[44:], np.int16).astype(np.float64)
    pcm_out = np.frombuffer(d[44:], np.int16).astype(np.float64)
    assert len(pcm_in) == len(pcm_out)
    snr = 10 * np.log10((pcm_in ** 2).sum() / ((pcm_in - pcm_out) ** 2).sum())
    assert snr > 15.0


def test_class_surface_and_errors(port):
    w = synth.wav(2, 2, 5000)
    assert HCA(w).encode() == port.hca_encode(w, 1)[1]
    assert HCA(w).encode(quality_level=CriHcaQuality.Lowest) == port.hca_encode(w, 1)[1]   # Lowest=5 falls back to High
    enc = HCA(w).encode(encrypt=True)
    assert enc == port.hca_crypt(port.hca_encode(w, 1)[1], 1, 56, 0xCF222F1FE0748978)[1]
    with pytest.raises(ValueError, match="must be a WAV"):
        HCA(port.hca_encode(w, 1)[1]).encode()


@pytest.mark.parametrize("quality", [0, 1, 2, 3])
def test_hard_material_matches_oracle(port, ctx, quality):
    """tests/helpers/hard_material.py (noise, full-scale squares, impulses, a step) at every quality, mono and stereo:
    byte-exact with the oracle, which tests/test_oracle_golden.py ties to the compiled reference on the same material."""
    from helpers import hard_material
    cases = hard_material.wavs(11 + quality)
    got = HCA.encode_batch(cases, Q[quality], ctx=ctx, raise_errors=False)
    for k, (w, g) in enumerate(zip(cases, got)):
        r, want = port.hca_encode(w, quality)
        if r == 0:
            assert not isinstance(g, Exception), (k, g)
            assert _frame_diff(g, want) is None, (k, quality)
        else:
            assert isinstance(g, Exception), (k, quality)
