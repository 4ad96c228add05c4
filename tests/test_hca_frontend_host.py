"""Host only: `pycricodecs_b200.HCA`'s header sniffing (a table-driven chunk walker) reports what the reference's class
reports -- same `info()` keys, order and values, same flags, same exceptions -- on v1.x / v2.0 / v3.0 streams, encrypted
streams, looping streams and WAVs. The reference's own `PyCriCodecs/hca.py` is imported from baseline/_ref when present."""
import glob
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_PKG = os.path.join(ROOT, "baseline", "_ref")


@pytest.fixture(scope="module")
def RefHCA():
    if not os.path.isdir(os.path.join(REF_PKG, "PyCriCodecs")):
        pytest.skip("baseline/_ref (pip install --target of the reference) is not present")
    before = set(sys.modules)
    sys.path.insert(0, REF_PKG)
    try:
        from PyCriCodecs.hca import HCA
    finally:
        sys.path.remove(REF_PKG)
        for name in set(sys.modules) - before:              # leave no `CriCodecs` / `PyCriCodecs` behind for other tests
            if name.split(".")[0] in ("CriCodecs", "PyCriCodecs"):
                del sys.modules[name]
    return HCA


def _cases(port):
    from helpers import hca3gen, wavgen
    from pycricodecs_b200 import synth
    cases = [open(f, "rb").read() for f in sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "*.hca")))]
    wav = synth.wav(3, 2, 5000)
    hca = port.hca_encode(wav, 1)[1]
    cases += [wav, wavgen.loop_wav(50, 2, 3200, 1000, 3000), hca, port.hca_crypt(hca, 1, 56, 0x1234)[1], port.hca_crypt(hca, 1, 1, 0)[1]]
    cases.append(hca3gen.stream(seed=3, version=0x0102, bands_per_hfr=0, min_res=1, dec=True, ath=1, total=100, base=60, stereo=40))
    loop = hca3gen.header(version=0x0200, frames=9, min_res=1)          # + a loop chunk in front of ciph
    at = loop.index(b"ciph")
    loop = loop[:at] + b"loop" + (2).to_bytes(4, "big") + (7).to_bytes(4, "big") + (128).to_bytes(2, "big") + (300).to_bytes(2, "big") + loop[at:-2]
    loop = loop[:6] + (len(loop) + 2).to_bytes(2, "big") + loop[8:]
    cases.append(loop + hca3gen.crc16_fast(loop).to_bytes(2, "big") + bytes(64))
    return cases, wav, hca


def test_info_matches_the_reference_class(port, RefHCA):
    from pycricodecs_b200 import HCA
    cases, _, _ = _cases(port)
    looping = 0
    for c in cases:
        a, b = HCA(c), RefHCA(c)
        assert a.info() == b.info() and list(a.info()) == list(b.info())
        assert (a.filetype, a.key, a.encrypted, a.looping) == (b.filetype, b.key, getattr(b, "encrypted", False), getattr(b, "looping", False))
        if a.filetype == "wav" and a.looping:
            assert (a.LoopCount, a.LoopStartSample, a.LoopEndSample) == (b.LoopCount, b.LoopStartSample, b.LoopEndSample)
        assert a.get_header() == b.get_header()
        looping += a.looping
    assert looping >= 2


def test_constructor_errors_match(port, RefHCA):
    from pycricodecs_b200 import HCA
    _, wav, hca = _cases(port)
    for args in ((hca, -1, 0), (hca, 1 << 70, 0), (hca, 0, -2), (hca, 0, 70000), (b"X" * 40, 0, 0), (hca, "CF222F1FE0748978", "12")):
        got = want = None
        try:
            got = HCA(*args).key
        except Exception as e:
            got = (type(e).__name__, str(e))
        try:
            want = RefHCA(*args).key
        except Exception as e:
            want = (type(e).__name__, str(e))
        assert got == want, args[1:]


def test_note_chunk_is_skipped():
    """A `note` chunk in front of the samples is stepped over (the reference seeks to an absolute offset there and then
    fails to find the data chunk: deliberate difference)."""
    from pycricodecs_b200 import HCA, synth
    wav = synth.wav(1, 1, 2000)
    note = b"note" + (12).to_bytes(4, "little") + bytes(12)
    with_note = wav[:36] + note + wav[36:]
    with_note = with_note[:4] + (len(with_note) - 8).to_bytes(4, "little") + with_note[8:]
    h = HCA(with_note)
    assert h.filetype == "wav" and h.info()["dataSize"] == 4000
