"""HCA v3.0 decode: extra scalefactors for the derived HFR scales, delta-coded intensities, the
v3.0 HFR rule and resolution-0 bands rebuilt by the stream-long noise generator (hca.cpp:1297-1307, 1353-1355, 1382-1424,
1602-1635, 1652-1661). The reference has no v3.0 encoder, so inputs are synthetic (tests/helpers/hca3gen.py) and the
expected values are the compiled reference's decodes (tools/make_golden_v3.py -> tests/golden/v3_*.hca, v3_digests.json).
Frames are sized so that no code is read from the last 32 bits of a frame, where the reference's reader shifts by a
negative count (hca.cpp:243-262)."""
import hashlib
import json
import os

import numpy as np
import pytest

from helpers import hca3gen

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
D = json.load(open(os.path.join(GOLD, "v3_digests.json")))
h = lambda b: hashlib.sha256(bytes(b)).hexdigest()[:16]

CASES = [   # wider than the committed fixtures; expected = oracle port (pinned by the fixtures and, when present, oracle/_ref)
    dict(seed=1, frame_size=4096),
    dict(seed=2, frame_size=2048, channels=1, stereo=0, base=60, bands_per_hfr=8),
    dict(seed=4, frame_size=4096, min_res=1),
    dict(seed=5, frame_size=4096, bands_per_hfr=0, base=50, stereo=40, total=90),
    dict(seed=6, frame_size=4096, max_res=9, frames=20),
    dict(seed=7, frame_size=4096, base=30, stereo=20, bands_per_hfr=6, total=120),
    dict(seed=8, frame_size=2048, channels=1, stereo=0, base=100, bands_per_hfr=2, total=126),
    dict(seed=9, frame_size=4096, min_res=2, max_res=12, sf_max=60),
    dict(seed=10, frame_size=4096, frames=37, delay=300),
    dict(seed=11, frame_size=4096, version=0x0200, min_res=1),
    dict(seed=12, frame_size=8192, channels=4, frames=6),                  # two primary / secondary pairs
    dict(seed=13, frame_size=6144, channels=3, frames=6, base=50, stereo=20, bands_per_hfr=5, total=110),
    # intensities cut short (the decoder keeps the previous frame's): across units of the general kernels, two pairs
    dict(seed=14, frame_size=4096, frames=40, kept=0.4),
    dict(seed=15, frame_size=4096, frames=40, kept=0.9, min_res=1),
    dict(seed=16, frame_size=8192, channels=4, frames=20, kept=0.5),
    dict(seed=17, frame_size=4096, frames=40, version=0x0200, min_res=1, kept=0.4),
]


def _fixture(name):
    return open(os.path.join(GOLD, name + ".hca"), "rb").read()


def test_generator_reproduces_the_fixtures():
    for name, e in D.items():
        assert h(hca3gen.stream(**e["args"])) == e["hca_sha"], name


def test_oracle_port_matches_the_reference_decodes(port):
    for name, e in D.items():
        s = _fixture(name)
        r, wav = port.hca_decode(s)
        assert r == 0 and len(wav) == e["wav_len"] and h(wav) == e["wav_sha"], name
        nch = s[12]
        for f in range(e["args"]["frames"]):                 # the codes end well before the frame does
            r, u = port.hca_unpack(s, 0, f, nch)
            assert r == 0 and u["bits"] < e["args"]["frame_size"] * 8 - 48, (name, f)
        if name.endswith("noise"):                           # the fixtures do exercise the generator
            r, u = port.hca_unpack(s, 0, 0, nch)
            assert ((u["res"] == 0) & (u["sf"] > 0)).any()


def test_oracle_port_matches_the_compiled_reference(port, ref):
    for kw in CASES:
        s = hca3gen.stream(**kw)
        r, wav = port.hca_decode(s)
        assert r == 0 and wav == ref.hca_decode(s), kw


@pytest.mark.gpu
def test_fixtures_decode_bit_exact(ctx):
    from pycricodecs_b200 import HCA
    names = list(D)
    got = HCA.decode_batch([_fixture(n) for n in names], ctx=ctx)
    assert [h(g) for g in got] == [D[n]["wav_sha"] for n in names]
    assert HCA(_fixture(names[0])).decode() == got[0]        # single-stream front-end


@pytest.mark.gpu
def test_v3_streams_match_the_oracle(ctx, port):
    from pycricodecs_b200 import HCA
    streams = [hca3gen.stream(**kw) for kw in CASES]
    got = HCA.decode_batch(streams, ctx=ctx)
    for kw, s, g in zip(CASES, streams, got):
        r, want = port.hca_decode(s)
        assert r == 0
        if g != want:
            a = np.frombuffer(g[44:], np.int16); b = np.frombuffer(want[44:], np.int16)
            bad = np.flatnonzero(a != b)
            raise AssertionError(f"{kw}: {bad.size} samples differ, first at {bad[:4]}")


@pytest.mark.gpu
def test_mixed_v2_and_v3_batch(ctx, port):
    """v3.0 streams send the whole job down the general kernels; the v2.0 streams beside them must not change."""
    from pycricodecs_b200 import HCA
    gold2 = open(os.path.join(GOLD, "s5_stereo_4800_high.hca"), "rb").read()
    low2 = open(os.path.join(GOLD, "s5_stereo_4800_low.hca"), "rb").read()
    v3 = hca3gen.stream(seed=21, frame_size=4096, frames=40)            # longer than one unit: state crosses units
    streams = [gold2, v3, low2, _fixture("v3_mono_hfr_noise")]
    got = HCA.decode_batch(streams, ctx=ctx)
    for s, g in zip(streams, got):
        assert g == port.hca_decode(s)[1]


@pytest.mark.gpu
@pytest.mark.parametrize("run", [None, "1", "7"])
def test_kept_intensities_on_the_fast_path(ctx, port, monkeypatch, run):
    """v2.0 pairs whose first intensity index is 15 keep the previous frame's other seven (hca.cpp:1368-1372): a batch of
    such streams alone takes the fast kernels, where the chain of kept values crosses run boundaries."""
    from pycricodecs_b200 import HCA
    if run:
        monkeypatch.setenv("CRI_HCA_FAST_RUN", run)
    streams = [hca3gen.stream(seed=40 + i, frame_size=3072, frames=50 + i, version=0x0200, min_res=1, kept=k)
               for i, k in enumerate((0.3, 0.95, 0.0, 0.6, 1.0))]
    streams.insert(2, _fixture("v2_joint_kept_intensities"))
    got = HCA.decode_batch(streams, ctx=ctx)
    for s, g in zip(streams, got):
        assert g == port.hca_decode(s)[1]
    assert h(got[2]) == D["v2_joint_kept_intensities"]["wav_sha"]


@pytest.mark.gpu
def test_unsupported_v3_layout_is_reported(ctx):
    """coded + 2 * hfr_groups >= 128: the reference's derived-scale copy reads a scalefactor left by the previous frame."""
    from pycricodecs_b200 import HCA
    s = hca3gen.stream(seed=30, frame_size=4096, channels=1, stereo=0, base=100, bands_per_hfr=1, total=128)
    ok = hca3gen.stream(seed=31, frame_size=4096)
    res = HCA.decode_batch([s, ok], ctx=ctx, raise_errors=False)
    assert isinstance(res[0], Exception) and res[0].status == -300
    assert isinstance(res[1], bytes)
