"""AWB batch front-end against a bank built, read and decoded by the reference's own
AWBBuilder / AWB / HCA classes (tools/make_golden_awb.py -> tests/golden/bank.awb, bank_digests.json)."""
import hashlib
import json
import os

import pytest

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
KEY = 0xCF222F1FE0748978
h = lambda b: hashlib.sha256(b).hexdigest()[:16]


def _bank():
    from pycricodecs_b200.awb import AWB
    return AWB(open(os.path.join(GOLD, "bank.awb"), "rb").read()), json.load(open(os.path.join(GOLD, "bank_digests.json")))


def test_header_and_files_match_the_reference_reader():
    a, d = _bank()
    assert (a.numfiles, a.align, a.subkey, a.version, a.headersize) == (d["numfiles"], d["align"], d["subkey"], d["version"], d["headersize"])
    assert a.ofs == d["ofs"]
    files = a.getfiles()
    assert [len(f) for f in files] == d["file_sizes"]
    assert [h(f) for f in files] == d["files"]
    assert h(a.getfile_atindex(2)) == d["files"][2]
    from pycricodecs_b200.awb import AWB
    assert AWB(os.path.join(GOLD, "bank.awb")).ofs == d["ofs"]          # path constructor
    with pytest.raises(ValueError, match="Invalid AWB header"):
        AWB(b"AFS3" + bytes(60))


@pytest.mark.gpu
def test_bank_decodes_in_one_batch_call(ctx, tmp_path, monkeypatch):
    a, d = _bank()
    launches = ctx.launches
    wavs = a.decode_all(KEY, ctx=ctx)
    assert ctx.launches - launches <= 4                                  # one job: patches + unpack + transform (+ clear)
    assert [None if w is None else h(w) for w in wavs] == d["wav"]
    monkeypatch.chdir(tmp_path)
    names = a.extract(decode=True, key=KEY, ctx=ctx)
    assert names == ["0.wav", "1.wav", "2.wav", "3.wav", "4.dat"]
    assert [h(open(n, "rb").read()) for n in names[:4]] == d["wav"][:4]
    assert a.extract(decode=False) == ["0.hca", "1.hca", "2.hca", "3.hca", "4.dat"]
