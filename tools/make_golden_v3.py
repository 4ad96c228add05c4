#!/usr/bin/env python3
"""Generate tests/golden/v3_*.hca and v3_digests.json: synthetic HCA v3.0 streams (tests/helpers/hca3gen.py) decoded by
the compiled REFERENCE (oracle/_ref). Runs only in the dev container."""
import hashlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle  # noqa: E402
from helpers import hca3gen  # noqa: E402

CASES = {   # name -> generator arguments
    "v3_joint_hfr_noise": dict(seed=101, frames=6, frame_size=3072),
    "v3_mono_hfr_noise": dict(seed=102, frames=6, frame_size=1536, channels=1, stereo=0, base=60, bands_per_hfr=8),
    "v3_joint_minres1": dict(seed=103, frames=5, frame_size=3072, min_res=1, base=30, stereo=20, bands_per_hfr=6, total=120),
    # intensities cut short: the decoder keeps the previous frame's values (hca.cpp:1410-1412 + :1185; v2.0: :1368-1372)
    "v3_joint_kept_intensities": dict(seed=104, frames=24, frame_size=3072, kept=0.5),
    "v2_joint_kept_intensities": dict(seed=105, frames=24, frame_size=3072, version=0x0200, min_res=1, kept=0.5),
}


def main():
    R = oracle.ref()
    gold = os.path.join(ROOT, "tests", "golden")
    d = {}
    for name, kw in CASES.items():
        s = hca3gen.stream(**kw)
        wav = R.hca_decode(s)
        open(os.path.join(gold, name + ".hca"), "wb").write(s)
        d[name] = {"args": kw, "hca_sha": hashlib.sha256(s).hexdigest()[:16], "wav_sha": hashlib.sha256(wav).hexdigest()[:16], "wav_len": len(wav)}
    json.dump(d, open(os.path.join(gold, "v3_digests.json"), "w"), indent=1)
    print(json.dumps(d, indent=1))


if __name__ == "__main__":
    main()
