#!/usr/bin/env python3
"""Generate tests/golden/bank.awb (+ bank_digests.json) with the REFERENCE's own code: streams encoded / encrypted by
the compiled reference (oracle/_ref), packed by the reference's `AWBBuilder`, read back and decoded one by one by the
reference's `AWB` + `HCA` classes (PyCriCodecs/awb.py, hca.py) running on top of the compiled reference module.
Runs only in the dev container (needs /root/reference and oracle/_ref)."""
import hashlib
import json
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
from pycricodecs_b200 import synth  # noqa: E402

KEY, SUBKEY = 0xCF222F1FE0748978, 0x1234
h = lambda b: hashlib.sha256(b).hexdigest()[:16]


def main():
    R = oracle.ref()
    sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref"))      # the compiled `CriCodecs` module
    sys.path.insert(0, "/root/reference")
    from PyCriCodecs.awb import AWB, AWBBuilder                     # the reference's own reader / builder
    from PyCriCodecs.hca import HCA
    out = os.path.join(ROOT, "tests", "golden")
    with tempfile.TemporaryDirectory() as tmp:
        names = []
        specs = [(11, 2, 3000, 1), (12, 1, 5000, 1), (13, 2, 1024 * 4 + 77, 0), (14, 2, 2500, 3)]   # sid, ch, samples, quality
        for i, (sid, ch, n, q) in enumerate(specs):
            r, x = R.hca_encode(synth.wav(sid, ch, n), q)
            assert r == 0
            x = R.hca_crypt(x, 1, 56, KEY, SUBKEY)
            p = os.path.join(tmp, f"{i}.hca")
            open(p, "wb").write(x)
            names.append(p)
        p = os.path.join(tmp, "4.dat")
        open(p, "wb").write(b"not an hca stream" * 3)
        names.append(p)
        bank = os.path.join(out, "bank.awb")
        AWBBuilder(names, subkey=SUBKEY, version=2, align=0x20).build(bank)
        a = AWB(bank)
        files = list(a.getfiles())
        d = {"numfiles": a.numfiles, "align": a.align, "subkey": a.subkey, "version": a.version, "ofs": a.ofs,
             "headersize": a.headersize, "files": [h(f) for f in files], "file_sizes": [len(f) for f in files], "wav": []}
        for f in files:
            d["wav"].append(h(HCA(f, key=KEY, subkey=a.subkey).decode()) if f[:4] in (b"HCA\x00", b"\xc8\xc3\xc1\x00") else None)
        json.dump(d, open(os.path.join(out, "bank_digests.json"), "w"), indent=1)
        print(d)


if __name__ == "__main__":
    main()
