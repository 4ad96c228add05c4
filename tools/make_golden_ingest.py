#!/usr/bin/env python3
"""tests/golden/ingest_digests.json: the COMPILED REFERENCE's ADX / HCA encodes of WAV images in every sample encoding
its loader accepts (tests/helpers/wavgen.py builds the images). Dev container only (needs oracle/_ref)."""
import hashlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle  # noqa: E402
from helpers import wavgen  # noqa: E402

h = lambda b: hashlib.sha256(b).hexdigest()[:16]


def main():
    R = oracle.ref()
    out = {}
    for kind, sid, ch, n in wavgen.CASES:
        w = wavgen.wav_as(kind, sid, ch, n)
        r, a = R.adx_encode(w)
        assert r == 0, (kind, r)
        r, x = R.hca_encode(w, 1)
        assert r == 0, (kind, r)
        out[kind] = {"wav": h(w), "adx": h(a), "adx_len": len(a), "hca": h(x), "hca_len": len(x)}
    loops = []
    for sid, ch, n, ls, le, ver in wavgen.LOOP_CASES:
        w = wavgen.loop_wav(sid, ch, n, ls, le)
        r, a = R.adx_encode(w, version=ver)
        assert r == 0, (sid, r)
        try:
            dec = h(R.adx_decode(a))                    # the reference may refuse its own output (copyright check quirk)
        except Exception as e:                          # noqa: BLE001
            dec = "error: " + str(e)
        loops.append({"wav": h(w), "adx": h(a), "adx_len": len(a), "adx_decoded": dec})
    out["adx_loops"] = loops
    hloops = []
    for sid, ch, n, ls, le, q in wavgen.HCA_LOOP_CASES:
        w = wavgen.loop_wav(sid, ch, n, ls, le)
        r, x = R.hca_encode(w, q)
        assert r == 0, (sid, r)
        hloops.append({"wav": h(w), "hca": h(x), "hca_len": len(x), "hca_decoded": h(R.hca_decode(x))})
    out["hca_loops"] = hloops
    json.dump(out, open(os.path.join(ROOT, "tests", "golden", "ingest_digests.json"), "w"), indent=1, sort_keys=True)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
