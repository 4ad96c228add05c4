"""GPU parity: ADX kernels (through the C-ABI) vs the oracle, bit-exact."""
import numpy as np
import pytest

from pycricodecs_b200 import engine, synth
from pycricodecs_b200.adx import ADX

pytestmark = pytest.mark.gpu


def _wav(sid, ch, n=synth.DEFAULT_SAMPLES):
    return synth.wav(sid, ch, n)


def test_encode_matches_oracle(port, ctx):
    wavs = [_wav(s, c) for s in range(6) for c in (1, 2)]
    got = ADX.encode_batch(wavs, ctx)
    for w, g in zip(wavs, got):
        r, want = port.adx_encode(w)
        assert r == 0 and g == want


def test_decode_matches_oracle(port, ctx):
    adx = [port.adx_encode(_wav(s, c))[1] for s in range(6) for c in (1, 2)]
    got = ADX.decode_batch(adx, ctx)
    for a, g in zip(adx, got):
        r, want = port.adx_decode(a)
        assert r == 0 and g == want


def test_round_trip_single_calls(port):
    w = _wav(3, 2, 32 * 50)
    a = ADX.encode(w)
    assert a == port.adx_encode(w)[1]
    assert ADX.decode(a) == port.adx_decode(a)[1]


@pytest.mark.parametrize("mode,depth,block", [(3, 4, 18), (4, 4, 18), (2, 4, 18), (3, 8, 18), (3, 2, 34), (4, 5, 12), (3, 15, 32)])
def test_modes_and_depths(port, ctx, mode, depth, block):
    wavs = [_wav(s, c, 2048) for s in (0, 5) for c in (1, 2)]
    enc = ADX.encode_batch(wavs, ctx, BitDepth=depth, Blocksize=block, Encoding=mode, Filter=2 if mode == 2 else 0)
    for w, g in zip(wavs, enc):
        r, want = port.adx_encode(w, depth, block, mode, 500, 2 if mode == 2 else 0, 4)
        assert r == 0 and g == want
    dec = ADX.decode_batch(enc, ctx, raise_errors=False)
    for a, g in zip(enc, dec):
        r, want = port.adx_decode(a)
        if r == 0:
            assert g == want
        else:
            assert isinstance(g, Exception) and g.status == r


def test_ragged_and_edge_lengths(port, ctx):
    # lengths that are not multiples of 32 exercise the reference's padding rule (adx.cpp:450-456)
    wavs = [_wav(s, c, n) for s, c, n in [(0, 1, 32), (1, 2, 33), (2, 1, 100), (3, 2, 1000), (4, 2, 31), (5, 1, 4096 + 16)]]
    enc = ADX.encode_batch(wavs, ctx)
    for w, g in zip(wavs, enc):
        assert g == port.adx_encode(w)[1]
    dec = ADX.decode_batch(enc, ctx, raise_errors=False)
    for a, g in zip(enc, dec):
        r, want = port.adx_decode(a)
        assert (g == want) if r == 0 else (g.status == r)


def test_many_channels_and_versions(port, ctx):
    n = 640
    pcm = np.stack([synth.channel_samples(9, c, n) for c in range(6)], axis=1)
    w6 = synth.wav_header(6, n) + pcm.tobytes()
    for ver in (3, 4, 5):
        g = ADX.encode_batch([w6], ctx, AdxVersion=ver)[0]
        assert g == port.adx_encode(w6, version=ver)[1]
        r, want = port.adx_decode(g)
        d = ADX.decode_batch([g], ctx, raise_errors=False)[0]
        assert (d == want) if r == 0 else (d.status == r)


def test_error_codes(ctx):
    w = _wav(0, 1, 64)
    with pytest.raises(ValueError, match="Blocksize"):
        ADX.encode(w, Blocksize=2)
    with pytest.raises(ValueError, match="Bitdepth"):
        ADX.encode(w, BitDepth=16)
    with pytest.raises(ValueError, match="Invalid ADX file header"):
        ADX.decode(b"\x00" * 64)
    bad = bytearray(ADX.encode(w)); bad[19] = 8
    with pytest.raises(NotImplementedError):
        ADX.decode(bytes(bad))


def test_eof_marker_stops_decode(port, ctx):
    a = bytearray(port.adx_encode(_wav(2, 2, 32 * 40))[1])
    hdr = int.from_bytes(a[2:4], "big") + 4
    a[hdr + 36 * 10: hdr + 36 * 10 + 2] = b"\x80\x01"   # EOF marker at frame 10
    got = ADX.decode_batch([bytes(a)], ctx)[0]
    assert got == port.adx_decode(bytes(a))[1]


def _loud_wav(seed, ch, n):
    """Full-scale material: square bursts, white noise at +-32767 and rail-to-rail steps. It drives the decoder's
    int16 clamp and the encoder's [-8, 7] delta clamp, i.e. the exact fallback behind the kernels' speculative
    (clamp-free) recurrences."""
    rng = np.random.default_rng(seed)
    x = rng.integers(-32768, 32768, size=(n, ch)).astype(np.int16)
    x[n // 4: n // 2] = np.where((np.arange(n // 4, n // 2) // 3 % 2)[:, None] == 0, 32767, -32768)
    x[n // 2: n // 2 + 64] = 0
    x[:64] = (x[:64].astype(np.int32) * np.arange(64)[:, None] // 64 // 64).astype(np.int16)   # quiet start: scale of block 0 < 256
    return synth.wav_header(ch, n) + x.tobytes()


def test_clipping_material_takes_the_exact_path(port, ctx):
    wavs = [_loud_wav(s, c, 32 * 200) for s in range(4) for c in (1, 2)]
    enc = ADX.encode_batch(wavs, ctx)
    for w, g in zip(wavs, enc):
        assert g == port.adx_encode(w)[1]
    dec = ADX.decode_batch(enc, ctx, raise_errors=False)
    for a, g in zip(enc, dec):
        r, want = port.adx_decode(a)
        assert (g == want) if r == 0 else (g.status == r)
    # a decoder-side stress the encoder never produces: huge scales on every block
    a = bytearray(enc[1])
    hdr = int.from_bytes(a[2:4], "big") + 4
    for b in range(hdr, len(a) - 18, 18):
        a[b] = 0x3F
    r, want = port.adx_decode(bytes(a))
    got = ADX.decode_batch([bytes(a)], ctx, raise_errors=False)[0]
    assert (got == want) if r == 0 else (got.status == r)
