#!/usr/bin/env python3
"""Generate tests/golden/v1_*.hca and v1_digests.json: synthetic HCA v1.x streams (`dec` chunk, ATH curve type 1;
tests/helpers/hca3gen.py) decoded by the compiled REFERENCE (oracle/_ref). The reference encoder only writes v2.0, so this
is the one way to pin the v1.x header path (hca.cpp:456-471, 710-727, 745-756). Runs only in the dev container."""
import hashlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle  # noqa: E402
from helpers import hca3gen  # noqa: E402

V1 = dict(bands_per_hfr=0, min_res=1, dec=True)
CASES = {   # name -> generator arguments
    # v1.01, joint stereo, no ath chunk -> ATH type 1 by default; 48 kHz: the curve runs into its 0xFF tail
    "v1_01_joint_default_ath": dict(seed=201, version=0x0101, frames=5, frame_size=3072, total=100, base=60, stereo=40, rate=48000, **V1),
    # v1.02, mono, explicit ath chunk type 0
    "v1_02_mono_ath0": dict(seed=202, version=0x0102, frames=5, frame_size=1536, channels=1, total=128, base=128, stereo=0, ath=0, **V1),
    # v1.03, discrete stereo (stereo type 0), explicit ath chunk type 1, 22.05 kHz
    "v1_03_discrete_ath1": dict(seed=203, version=0x0103, frames=5, frame_size=3072, total=112, base=112, stereo=0, ath=1, rate=22050, **V1),
    # three channels (pair + discrete): the general kernels
    "v1_01_three_channels": dict(seed=204, version=0x0101, frames=4, frame_size=4608, channels=3, total=96, base=64, stereo=32, rate=32000, **V1),
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
    json.dump(d, open(os.path.join(gold, "v1_digests.json"), "w"), indent=1)
    print(json.dumps(d, indent=1))


if __name__ == "__main__":
    main()
