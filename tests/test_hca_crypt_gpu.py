"""GPU parity: HCA en/decryption (HcaCrypt) through the C-ABI vs the oracle, byte-exact."""
import pytest

from pycricodecs_b200 import HCA, synth

pytestmark = pytest.mark.gpu
KEY = 0xCF222F1FE0748978


def _plain(port, sid, ch=2, n=30000, q=1):
    return port.hca_encode(synth.wav(sid, ch, n), q)[1]


def test_encrypt_matches_oracle(port, ctx):
    plain = [_plain(port, s, 1 + s % 2, 20000 + 777 * s, s % 4) for s in range(6)]
    got = HCA.crypt_batch(plain, True, keys=KEY, ctx=ctx)
    for p, g in zip(plain, got):
        assert g == port.hca_crypt(p, 1, 56, KEY)[1]


def test_decrypt_round_trip_and_subkey(port, ctx):
    plain = [_plain(port, s) for s in range(4)]
    keys = [KEY, 0x1234567, KEY, 1]
    subs = [0, 0, 0x1234, 0xFFFF]
    enc = HCA.crypt_batch(plain, True, keys=keys, subkeys=subs, ctx=ctx)
    for p, e, k, s in zip(plain, enc, keys, subs):
        assert e == port.hca_crypt(p, 1, 56, k, s)[1]
    dec = HCA.crypt_batch(enc, False, keys=keys, subkeys=subs, ctx=ctx)
    assert dec == plain


def test_keyless_type1(port, ctx):
    p = _plain(port, 9)
    e = HCA.crypt_batch([p], True, keyless=True, ctx=ctx)[0]
    assert e == port.hca_crypt(p, 1, 1, 0)[1]
    assert HCA.crypt_batch([e], False, ctx=ctx)[0] == p


def test_class_encrypt_decrypt(port):
    p = _plain(port, 3)
    h = HCA(p)
    h.encrypt(KEY)
    assert h.get_hca() == port.hca_crypt(p, 1, 56, KEY)[1]
    h2 = HCA(h.get_hca(), key=KEY)
    assert h2.encrypted and h2.decode() == port.hca_decode(p)[1]
    h2.decrypt(KEY)
    assert h2.get_hca() == p
    with pytest.raises(ValueError, match="already decrypted"):
        h2.decrypt(KEY)


def test_bad_header_status(ctx):
    res = HCA.crypt_batch([b"\x00" * 200], True, keys=KEY, ctx=ctx, raise_errors=False)
    assert res[0].status == -201
