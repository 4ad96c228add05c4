"""HCA v1.x decode: the `dec` chunk (hca.cpp:710-727), the ATH curve of type 1 and its default
below v2.0 (hca.cpp:456-471, 745-756). The reference encoder only writes v2.0, so inputs are synthetic
(tests/helpers/hca3gen.py with dec=True) and the expected values are the compiled reference's decodes
(tools/make_golden_v1.py -> tests/golden/v1_*.hca, v1_digests.json)."""
import hashlib
import json
import os

import numpy as np
import pytest

from helpers import hca3gen

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
D = json.load(open(os.path.join(GOLD, "v1_digests.json")))
h = lambda b: hashlib.sha256(bytes(b)).hexdigest()[:16]
V1 = dict(bands_per_hfr=0, min_res=1, dec=True)

CASES = [   # wider than the committed fixtures; expected = oracle port (pinned by the fixtures and, when present, oracle/_ref)
    dict(seed=211, version=0x0101, frame_size=3072, total=128, base=100, stereo=28, rate=44100, **V1),
    dict(seed=212, version=0x0101, frame_size=1536, channels=1, total=90, base=90, stereo=0, rate=8000, **V1),
    dict(seed=213, version=0x0102, frame_size=3072, total=128, base=128, stereo=0, ath=1, rate=96000, **V1),
    dict(seed=214, version=0x0103, frame_size=3072, total=64, base=40, stereo=24, ath=0, frames=20, **V1),
    dict(seed=215, version=0x0101, frame_size=6144, channels=4, total=100, base=70, stereo=30, rate=48000, frames=6, **V1),
    dict(seed=216, version=0x0101, frame_size=3072, total=100, base=60, stereo=40, rate=48000, dec=False, bands_per_hfr=0, min_res=1),  # comp chunk, v1 defaults
]


def _fixture(name):
    return open(os.path.join(GOLD, name + ".hca"), "rb").read()


def test_generator_reproduces_the_fixtures():
    for name, e in D.items():
        assert h(hca3gen.stream(**e["args"])) == e["hca_sha"], name


def test_oracle_port_matches_the_reference_decodes(port):
    for name, e in D.items():
        r, wav = port.hca_decode(_fixture(name))
        assert r == 0 and len(wav) == e["wav_len"] and h(wav) == e["wav_sha"], name


def test_the_ath_curve_is_exercised(port):
    """Same frames with and without the curve decode differently (the curve moves the resolutions), and the 48 kHz curve
    reaches its 0xFF tail (hca.cpp:462-466)."""
    kw = dict(D["v1_01_joint_default_ath"]["args"])
    with_curve = hca3gen.stream(**kw)
    without = hca3gen.stream(**dict(kw, ath=0))
    assert with_curve[-3072 * 5:] == without[-3072 * 5:]                    # identical frames, only the header differs
    a, b = port.hca_decode(with_curve)[1], port.hca_decode(without)[1]
    assert len(a) == len(b) and a != b
    r, info = port.hca_info(with_curve)
    assert r == 0


def test_oracle_port_matches_the_compiled_reference(port, ref):
    for kw in CASES:
        s = hca3gen.stream(**kw)
        r, wav = port.hca_decode(s)
        assert r == 0 and wav == ref.hca_decode(s), kw


@pytest.mark.gpu
def test_fixtures_decode_bit_exact(ctx):
    from pycricodecs_b200 import HCA
    names = list(D)
    got = HCA.decode_batch([_fixture(n) for n in names], ctx=ctx)        # mixed channel counts: the general kernels
    assert [h(g) for g in got] == [D[n]["wav_sha"] for n in names]
    for n in names:                                                        # one at a time: mono / stereo take the fast kernels
        assert h(HCA.decode_batch([_fixture(n)], ctx=ctx)[0]) == D[n]["wav_sha"], n
    assert HCA(_fixture(names[0])).decode() == got[0]
    assert HCA(_fixture(names[0])).info()["version"] == "0x101"


@pytest.mark.gpu
def test_v1_streams_match_the_oracle(ctx, port):
    from pycricodecs_b200 import HCA
    streams = [hca3gen.stream(**kw) for kw in CASES]
    stereo = [s for s in streams if s[12] == 2]
    for batch in (streams, stereo):                                        # general kernels, then the fast path (stereo only)
        got = HCA.decode_batch(batch, ctx=ctx)
        for s, g in zip(batch, got):
            r, want = port.hca_decode(s)
            assert r == 0
            if g != want:
                a = np.frombuffer(g[44:], np.int16); b = np.frombuffer(want[44:], np.int16)
                bad = np.flatnonzero(a != b)
                raise AssertionError(f"{bad.size} samples differ, first at {bad[:4]}")
