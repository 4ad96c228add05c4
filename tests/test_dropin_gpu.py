"""GPU: the `CriCodecs`-named drop-in module has the reference's call signatures and results."""
import importlib.util
import os

import pytest

from pycricodecs_b200 import synth

pytestmark = pytest.mark.gpu
KEY = 0xCF222F1FE0748978


@pytest.fixture(scope="module")
def mod():
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "pycricodecs_b200", "dropin", "CriCodecs.py")
    spec = importlib.util.spec_from_file_location("CriCodecs_dropin", path)
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def test_five_callables(port, mod):
    w = synth.wav(4, 2, 6400)
    a = mod.AdxEncode(w, 4, 18, 3, 500, 0, 4, False)
    assert a == port.adx_encode(w)[1]
    assert mod.AdxDecode(a) == port.adx_decode(a)[1]
    h = mod.HcaEncode(w, 0, 1)
    assert h == port.hca_encode(w, 1)[1]
    e = mod.HcaCrypt(bytearray(h), 1, 96, 56, KEY, 0)
    assert e == port.hca_crypt(h, 1, 56, KEY)[1]
    assert mod.HcaCrypt(e, 0, 96, 0, KEY, 0) == h
    assert mod.HcaDecode(e, 96, KEY, 0) == port.hca_decode(h)[1]
    with pytest.raises(ValueError, match="Blocksize"):
        mod.AdxEncode(w, 4, 2, 3, 500, 0, 4, False)
