"""GPU parity: HCA decode kernels (through the C-ABI) vs the oracle, bit-exact PCM."""
import numpy as np
import pytest

from pycricodecs_b200 import HCA, synth

pytestmark = pytest.mark.gpu
KEY = 0xCF222F1FE0748978


def _first_diff(a: bytes, b: bytes):
    x = np.frombuffer(a, np.uint8); y = np.frombuffer(b, np.uint8)
    if len(x) != len(y):
        return f"len {len(x)} vs {len(y)}"
    d = np.nonzero(x != y)[0]
    return None if len(d) == 0 else f"{len(d)} bytes differ, first at {d[0]} (frame {(d[0] - 44) // 4096 if d[0] >= 44 else 'hdr'})"


@pytest.mark.parametrize("quality", [0, 1, 2, 3])
@pytest.mark.parametrize("channels", [1, 2])
def test_decode_matches_oracle(port, ctx, quality, channels):
    hcas = [port.hca_encode(synth.wav(s, channels), quality)[1] for s in range(3)]
    got = HCA.decode_batch(hcas, ctx=ctx)
    for h, g in zip(hcas, got):
        r, want = port.hca_decode(h)
        assert r == 0
        assert _first_diff(g, want) is None


def test_short_and_ragged_streams(port, ctx):
    cases = [(0, 2, 1), (1, 2, 127), (2, 1, 128), (3, 2, 129), (4, 2, 1024), (5, 1, 1024 - 128), (6, 2, 1024 * 13 + 5), (7, 2, 40000)]
    hcas = [port.hca_encode(synth.wav(s, c, n), 1)[1] for s, c, n in cases]
    got = HCA.decode_batch(hcas, ctx=ctx)
    for h, g in zip(hcas, got):
        assert _first_diff(g, port.hca_decode(h)[1]) is None


def test_encrypted_decode(port, ctx):
    plain = [port.hca_encode(synth.wav(s, 2, 20000), q)[1] for s, q in [(0, 1), (1, 3)]]
    enc = [port.hca_crypt(p, 1, 56, KEY)[1] for p in plain] + [port.hca_crypt(plain[0], 1, 1, 0)[1], port.hca_crypt(plain[1], 1, 56, KEY, 0x1234)[1]]
    got = HCA.decode_batch(enc, keys=[KEY, KEY, 0, KEY], subkeys=[0, 0, 0, 0x1234], ctx=ctx)
    want = [port.hca_decode(plain[0])[1], port.hca_decode(plain[1])[1], port.hca_decode(plain[0])[1], port.hca_decode(plain[1])[1]]
    for g, w in zip(got, want):
        assert _first_diff(g, w) is None


def test_corrupt_frame_reports_decode_error(port, ctx):
    h = bytearray(port.hca_encode(synth.wav(0, 2, 8192), 1)[1])
    h[96 + 682 * 3 + 100] ^= 0x55          # breaks the CRC of frame 3
    res = HCA.decode_batch([bytes(h), b"junk" * 30], ctx=ctx, raise_errors=False)
    assert isinstance(res[0], Exception) and res[0].status == -202
    assert isinstance(res[1], Exception) and res[1].status == -201
    with pytest.raises(ValueError, match="Decoding error"):
        HCA(bytes(h)).decode()


def test_class_surface(port):
    h = port.hca_encode(synth.wav(2, 2, 5000), 1)[1]
    obj = HCA(h)
    assert obj.info()["FrameCount"] == 6 and obj.filetype == "hca"
    assert obj.decode() == port.hca_decode(h)[1]


@pytest.mark.parametrize("run_len", [1, 2, 3, 5, 7, 64])
@pytest.mark.parametrize("channels", [1, 2])
def test_fast_path_run_lengths(port, ctx, monkeypatch, run_len, channels):
    """The fast path cuts the flattened frame list into runs: force run lengths that make runs start in the middle
    of streams, cross from one stream into the next and end short, with more than one transform warp of runs."""
    monkeypatch.setenv("CRI_HCA_FAST_RUN", str(run_len))
    lengths = [5000, 1, 1024 * 3, 1024 * 9 + 17, 700, 1024 * 20, 2047, 1024 * 6 - 128, 31000, 900]
    hcas = [port.hca_encode(synth.wav(s, channels, n), s % 2)[1] for s, n in enumerate(lengths)]   # Highest and High
    hcas = hcas * 3
    got = HCA.decode_batch(hcas, ctx=ctx)
    want = {h: port.hca_decode(h)[1] for h in set(hcas)}
    for h, g in zip(hcas, got):
        assert _first_diff(g, want[h]) is None


def test_fast_and_general_path_agree(port, ctx, monkeypatch):
    hcas = [port.hca_encode(synth.wav(s, 2, 1024 * 11 + 3 * s), 1)[1] for s in range(5)]
    fast = HCA.decode_batch(hcas, ctx=ctx)
    monkeypatch.setenv("CRI_HCA_GENERAL", "1")
    general = HCA.decode_batch(hcas, ctx=ctx)
    assert fast == general


def test_fast_path_bad_frame_in_the_middle_of_a_batch(port, ctx):
    good = [port.hca_encode(synth.wav(s, 2, 9000), 1)[1] for s in range(4)]
    bad = bytearray(good[1]); bad[96 + 682 * 2 + 50] ^= 0x10
    res = HCA.decode_batch([good[0], bytes(bad), good[2], good[3]], ctx=ctx, raise_errors=False)
    assert isinstance(res[1], Exception) and res[1].status == -202
    for i in (0, 2, 3):
        assert res[i] == port.hca_decode(good[i])[1]


@pytest.mark.parametrize("run_len", [0, 3, 8])
@pytest.mark.parametrize("channels", [1, 2])
def test_fast_path_mixed_qualities(port, ctx, monkeypatch, run_len, channels):
    """One batch with discrete-channel streams (Highest, High), intensity-stereo streams (Middle) and intensity + HFR
    streams (Low): all of it is on the fast path (JOINT kernels), with lanes of a warp on different kinds of stream."""
    if run_len:
        monkeypatch.setenv("CRI_HCA_FAST_RUN", str(run_len))
    lengths = [4000, 1024 * 5 + 1, 9000, 1500, 1024 * 12, 333, 7777, 1024 * 3]
    hcas = [port.hca_encode(synth.wav(20 + s, channels, n), s % 4)[1] for s, n in enumerate(lengths)] * 2
    got = HCA.decode_batch(hcas, ctx=ctx)
    want = {h: port.hca_decode(h)[1] for h in set(hcas)}
    for i, (h, g) in enumerate(zip(hcas, got)):
        assert _first_diff(g, want[h]) is None, f"stream {i} (quality {i % 4})"
    monkeypatch.setenv("CRI_HCA_GENERAL", "1")
    assert HCA.decode_batch(hcas, ctx=ctx) == got


def test_checked_reader_variants_agree(port, ctx, monkeypatch):
    """The unpack kernel has two variants of its header parser and of its code loop: one without per-read end-of-frame
    checks (used when a frame is long enough that no read can reach its end) and one with the reference reader's rule
    (a read that would cross the end yields 0, hca.cpp:232-233). Valid streams never need the second; force it."""
    hcas = [port.hca_encode(synth.wav(30 + s, 1 + s % 2, 6000 + 500 * s), s % 4)[1] for s in range(8)]
    want = HCA.decode_batch(hcas, ctx=ctx)
    monkeypatch.setenv("CRI_HCA_CAREFUL", "1")
    assert HCA.decode_batch(hcas, ctx=ctx) == want
    for h, g in zip(hcas[:4], want[:4]):
        assert g == port.hca_decode(h)[1]


def test_truncated_last_frame_matches_reference_reader(port, ctx):
    """A frame whose CRC is valid but whose code runs reach past its end: reads past the end are zeros. Built by
    shrinking frame_size in the header and re-stamping every frame's CRC; the oracle decodes the same bytes."""
    import struct
    from pycricodecs_b200 import engine
    src = port.hca_encode(synth.wav(5, 2, 4096), 0)[1]                     # Highest: 1024-byte frames, dense
    fs_old = struct.unpack(">H", src[28:30])[0]
    fs_new = fs_old - 300
    hdr = bytearray(src[:96])
    hdr[28:30] = struct.pack(">H", fs_new)
    hdr[94:96] = struct.pack(">H", engine.crc16(bytes(hdr[:94])))
    frames = []
    for f in range(struct.unpack(">I", src[16:20])[0]):
        body = bytearray(src[96 + f * fs_old: 96 + f * fs_old + fs_new - 2])
        frames.append(bytes(body) + struct.pack(">H", engine.crc16(bytes(body))))
    cut = bytes(hdr) + b"".join(frames)
    r, want = port.hca_decode(cut)
    got = HCA.decode_batch([cut], ctx=ctx, raise_errors=False)[0]
    assert (got == want) if r == 0 else (got.status == -202)
