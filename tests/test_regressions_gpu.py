"""GPU: regressions for the round-1 review findings -- a stream's result must not depend on its neighbours in the batch
(PCM16 on an odd byte of the blob), v2.0-layout streams with more HFR groups than the fast unpack kernel keeps in one word,
corrupt frames whose codes run far past the end of the frame."""
import numpy as np
import pytest

from helpers import hca3gen, wavgen

pytestmark = pytest.mark.gpu


def test_encode_does_not_depend_on_batch_neighbours(ctx, port):
    """An odd-sized WAV (8-bit mono, odd sample count) in front shifts the next PCM16 stream onto an odd byte."""
    from pycricodecs_b200 import engine, synth
    odd = wavgen.wav_as("u8", 90, 1, 32 * 40 + 1)
    assert len(odd) % 2 == 1
    wav = synth.wav(91, 2, 5000)
    alone_hca = engine.hca_encode_batch([wav], ctx=ctx)[0]
    alone_adx = engine.adx_encode_batch([wav], ctx=ctx)[0]
    assert alone_hca == port.hca_encode(wav, 1)[1] and alone_adx == port.adx_encode(wav)[1]
    got_hca = engine.hca_encode_batch([odd, wav, odd, wav], ctx=ctx, raise_errors=False)
    got_adx = engine.adx_encode_batch([odd, wav, odd, wav], ctx=ctx, raise_errors=False)
    assert got_hca[1] == alone_hca and got_hca[3] == alone_hca
    assert got_adx[1] == alone_adx and got_adx[3] == alone_adx


def test_many_hfr_groups_take_the_general_kernels(ctx, port):
    """bands_per_hfr = 1 with 20 reconstructed bands: 20 HFR scales per channel (v2.0 layout)."""
    from pycricodecs_b200 import HCA
    streams = [hca3gen.stream(seed=40 + k, version=0x0200, min_res=1, frame_size=4096, total=100, base=60, stereo=20, bands_per_hfr=1, frames=5)
               for k in range(3)]
    streams.append(hca3gen.stream(seed=44, version=0x0200, min_res=1, frame_size=2048, channels=1, total=120, base=90, stereo=0, bands_per_hfr=2, frames=5))
    for batch in (streams[:3], streams[3:]):
        got = HCA.decode_batch(batch, ctx=ctx)
        for s, g in zip(batch, got):
            r, want = port.hca_decode(s)
            assert r == 0 and g == want


def test_codes_running_far_past_a_short_frame(ctx, port):
    """A valid CRC over a tiny frame whose header promises 128 bands in 2 channels: every code lies beyond the frame. The
    readers return zeros there (hca.cpp:232-233) and must not fault, whatever lies behind the scratch rows."""
    from pycricodecs_b200 import HCA
    s = hca3gen.stream(seed=60, version=0x0200, min_res=1, frame_size=40, total=128, base=128, stereo=0, bands_per_hfr=0, frames=3, level=(300, 400))
    s3 = hca3gen.stream(seed=51, version=0x0200, min_res=1, frame_size=40, channels=3, total=128, base=128, stereo=0, bands_per_hfr=0, frames=3, level=(300, 400))
    for st, nch in ((s, 2), (s3, 3)):
        r, u = port.hca_unpack(st, 0, 0, nch)
        assert r == 0 and u["bits"] > 3 * 40 * 8                      # the codes of a frame end far behind the frame
    for batch in ([s] * 300, [s3] * 100 + [s] * 50):                   # fast kernels; general kernels (mixed channel counts)
        res = HCA.decode_batch(batch, ctx=ctx, raise_errors=False)
        want = {id(s): port.hca_decode(s), id(s3): port.hca_decode(s3)}
        for st, g in zip(batch, res):
            r, w = want[id(st)]
            assert r == 0 and g == w
