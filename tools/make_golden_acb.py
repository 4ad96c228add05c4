#!/usr/bin/env python3
"""Generate tests/golden/sheet.acb, sheet_ext.acb, sheet_masked.utf and sheet_digests.json with the REFERENCE's own code: the cue
sheet is built by the reference's `UTFBuilder` around tests/golden/bank.awb (tools/make_golden_awb.py), read back by
the reference's `UTF` / `ACB` classes and extracted with `ACB.extract(decode=True)` (PyCriCodecs/acb.py:141-154) on top
of the compiled reference module. Runs only in the dev container (needs /root/reference and oracle/_ref)."""
import hashlib
import json
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402

KEY = 0xCF222F1FE0748978
h = lambda b: hashlib.sha256(bytes(b)).hexdigest()[:16]


def plain(payload):
    """JSON-able view of a payload: type name + value (bytes as digest, nested tables recursively)."""
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


def plain_table(table):
    """JSON-able view of the reader's column-major `table` (bytes as digest, tuples as lists)."""
    conv = lambda v: {"len": len(v), "sha": h(v)} if isinstance(v, (bytes, bytearray)) else list(v) if isinstance(v, tuple) else v
    return {k: [conv(v) for v in col] for k, col in table.items()}


def main():
    oracle.ref()
    sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref"))
    sys.path.insert(0, "/root/reference")
    from PyCriCodecs.acb import ACB
    from PyCriCodecs.chunk import UTFTypeValues as T
    from PyCriCodecs.utf import UTF, UTFBuilder
    gold = os.path.join(ROOT, "tests", "golden")
    bank = open(os.path.join(gold, "bank.awb"), "rb").read()
    enc = [2, 6, 2, 2, 0]                               # HCA, HCA-MX, HCA, HCA, ADX
    waves = [{"MemoryAwbId": (T.ushort, i), "EncodeType": (T.uchar, e), "Streaming": (T.uchar, 0), "NumChannels": (T.uchar, 2 - (i == 1)),
              "SamplingRate": (T.ushort, 48000), "NumSamples": (T.uint, 1000 + i), "ExtensionData": (T.ushort, 0xFFFF)}
             for i, e in enumerate(enc)]
    wave_table = bytes(UTFBuilder(waves, table_name="Waveform").parse())
    cues = [{"CueName": (T.string, f"cue_{i}"), "CueIndex": (T.ushort, i)} for i in range(5)]
    cue_table = bytes(UTFBuilder(cues, table_name="CueName").parse())
    head = [{"FileIdentifier": (T.uint, 0), "Size": (T.uint, 0), "Version": (T.uint, 0x01290000), "Type": (T.uchar, 0),
             "VersionString": (T.string, "\nACB Format/PC Ver.1.29.0 Build:\n"), "Name": (T.string, "sheet"), "CategoryExtension": (T.uchar, 0),
             "AcbVolume": (T.float, 1.0), "AcbGuid": (T.bytes, bytes(range(16))), "WaveformTable": (T.bytes, wave_table),
             "CueNameTable": (T.bytes, cue_table), "AwbFile": (T.bytes, bank), "StreamAwbHash": (T.bytes, None)}]
    sheet = bytes(UTFBuilder(head, table_name="Header").parse())
    open(os.path.join(gold, "sheet.acb"), "wb").write(sheet)
    ext = [dict(head[0], AwbFile=(T.bytes, b""))]                 # external bank: <Name>.awb beside the sheet (acb.py:37-42)
    open(os.path.join(gold, "sheet_ext.acb"), "wb").write(bytes(UTFBuilder(ext, table_name="Header").parse()))
    masked = bytes(UTFBuilder(waves, encrypt=True, table_name="Waveform").parse())
    open(os.path.join(gold, "sheet_masked.utf"), "wb").write(masked)

    a = ACB(sheet)
    d = {"payload": plain(a.payload), "masked_payload": plain(UTF(masked).get_payload()),
         "table": plain_table(UTF(sheet).table), "masked_table": plain_table(UTF(masked).table), "cue_table": plain_table(UTF(cue_table).table),
         "table_name": UTF(sheet).table_name, "awb_numfiles": a.awb.numfiles, "awb_subkey": a.awb.subkey}
    with tempfile.TemporaryDirectory() as tmp:
        for mode in (True, False):
            sub = os.path.join(tmp, "d" if mode else "r")
            ACB(sheet).extract(decode=mode, key=KEY, dirname=sub)      # fresh reader: the reference's getfiles() is one-shot
            names = sorted(os.listdir(sub))
            d["decoded" if mode else "raw"] = {n: h(open(os.path.join(sub, n), "rb").read()) for n in names}
    json.dump(d, open(os.path.join(gold, "sheet_digests.json"), "w"), indent=1)
    print(json.dumps(d, indent=1)[:3000])


if __name__ == "__main__":
    main()
