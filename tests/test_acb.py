"""ACB cue-sheet front-end against a sheet built by the reference's UTFBuilder, read by the
reference's UTF / ACB classes and extracted by `ACB.extract` (tools/make_golden_acb.py -> tests/golden/sheet.acb,
sheet_masked.utf, sheet_digests.json)."""
import hashlib
import json
import os

import pytest

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
KEY = 0xCF222F1FE0748978
h = lambda b: hashlib.sha256(bytes(b)).hexdigest()[:16]
D = json.load(open(os.path.join(GOLD, "sheet_digests.json")))


def plain(payload):
    out = []
    for row in payload:
        r = {}
        for k, v in row.items():
            if isinstance(v, list):
                r[k] = ["table", plain(v)]
            elif isinstance(v[1], (bytes, bytearray)):
                r[k] = [v[0].name, {"len": len(v[1]), "sha": h(v[1])}]
            else:
                r[k] = [v[0].name, v[1]]
        out.append(r)
    return out


def test_utf_payload_matches_the_reference_reader():
    from pycricodecs_b200.acb import ACB
    from pycricodecs_b200.utf import UTF
    path = os.path.join(GOLD, "sheet.acb")
    a = ACB(path)                                                     # path constructor
    assert plain(a.payload) == D["payload"]
    assert [list(r) for r in plain(a.payload)] == [list(r) for r in D["payload"]]      # column order too
    assert UTF(path).table_name == D["table_name"]
    assert (a.awb.numfiles, a.awb.subkey) == (D["awb_numfiles"], D["awb_subkey"])
    b = ACB(open(path, "rb").read())                                  # bytes constructor
    assert plain(b.payload) == D["payload"]
    t = UTF(open(path, "rb").read())
    assert t.table["Name"] == ["sheet"] and t.num_rows == 1


def plain_table(table):
    conv = lambda v: {"len": len(v), "sha": h(v)} if isinstance(v, (bytes, bytearray)) else list(v) if isinstance(v, tuple) else v
    return {k: [conv(v) for v in col] for k, col in table.items()}


def test_utf_table_view_follows_the_reference_rules():
    """`UTF.table` (column-major): columns without storage hold one 0 / "<NULL>" / b"", constants one entry (numbers as
    the raw 1-tuple), per-row columns one entry per row, in that key order (PyCriCodecs/utf.py:113-152)."""
    from pycricodecs_b200.utf import UTF
    sheet = open(os.path.join(GOLD, "sheet.acb"), "rb").read()
    t = UTF(sheet)
    assert plain_table(t.table) == D["table"] and list(t.table) == list(D["table"])
    m = UTF(open(os.path.join(GOLD, "sheet_masked.utf"), "rb").read())
    assert plain_table(m.table) == D["masked_table"] and list(m.table) == list(D["masked_table"])
    cue = UTF(t.get_payload()[0]["CueNameTable"][1])
    assert plain_table(cue.table) == D["cue_table"] and list(cue.table) == list(D["cue_table"])


def test_masked_utf_is_unmasked():
    from pycricodecs_b200.utf import UTF
    t = UTF(open(os.path.join(GOLD, "sheet_masked.utf"), "rb").read())
    assert plain(t.get_payload()) == D["masked_payload"]
    assert t.table["EncodeType"] == [2, 6, 2, 2, 0]


def test_utf_rejects_garbage():
    from pycricodecs_b200.utf import UTF
    with pytest.raises(ValueError, match="UTF chunk is not present"):
        UTF(b"@UTG" + bytes(64))
    good = open(os.path.join(GOLD, "sheet.acb"), "rb").read()
    with pytest.raises(ValueError):
        UTF(good[:40])


def test_extract_without_decoding_matches_the_reference(tmp_path):
    from pycricodecs_b200.acb import ACB
    a = ACB(os.path.join(GOLD, "sheet.acb"))
    names = a.extract(decode=False, dirname=str(tmp_path / "raw"))
    assert [os.path.basename(n) for n in names] == sorted(D["raw"])
    assert {os.path.basename(n): h(open(n, "rb").read()) for n in names} == D["raw"]
    assert a.get_extension(0) == ".adx" and a.get_extension(6) == ".hca" and a.get_extension(9) is None


def test_external_bank_is_found_beside_the_sheet(tmp_path):
    import shutil
    from pycricodecs_b200.acb import ACB
    shutil.copy(os.path.join(GOLD, "sheet_ext.acb"), tmp_path / "x.acb")
    with pytest.raises(FileNotFoundError):
        ACB(str(tmp_path / "x.acb"))
    shutil.copy(os.path.join(GOLD, "bank.awb"), tmp_path / "sheet.awb")           # the sheet's Name is "sheet"
    a = ACB(str(tmp_path / "x.acb"))
    assert a.payload[0]["AwbFile"][1] == b"" and a.awb.numfiles == D["awb_numfiles"]
    names = a.extract(decode=False, dirname=str(tmp_path / "raw"))
    assert {os.path.basename(n): h(open(n, "rb").read()) for n in names} == D["raw"]


@pytest.mark.gpu
def test_sheet_decodes_in_one_batch_call(ctx, tmp_path):
    from pycricodecs_b200.acb import ACB
    a = ACB(os.path.join(GOLD, "sheet.acb"))
    launches = ctx.launches
    names = a.extract(decode=True, key=KEY, dirname=str(tmp_path / "wav"), ctx=ctx)
    assert ctx.launches - launches <= 4
    assert {os.path.basename(n): h(open(n, "rb").read()) for n in names} == D["decoded"]
    wavs = a.decode_all(KEY, ctx=ctx)
    assert [None if w is None else h(w) for w in wavs] == [D["decoded"].get(f"{i}.wav") for i in range(5)]
