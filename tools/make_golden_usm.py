#!/usr/bin/env python3
"""Generate tests/golden/usm_audio.json with the REFERENCE's own USM class (PyCriCodecs/usm.py:47-118, 313-322): key
schedule and AudioMask outputs for a few keys / payload sizes. Runs only in the dev container."""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402


def main():
    oracle.ref()
    sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref"))
    sys.path.insert(0, "/root/reference")
    from PyCriCodecs.usm import USM
    out = {"keys": [], "masked": []}
    for key in (0, 1, 0xCF222F1FE0748978, 0x0123456789ABCDEF, "1a2b3c", "FFFFFFFFFFFFFFFF"):
        u = USM.__new__(USM)
        u.init_key(key)
        out["keys"].append({"key": key, "videomask1": bytes(u.videomask1).hex(), "videomask2": bytes(u.videomask2).hex(),
                            "audiomask": bytes(u.audiomask).hex()})
    u = USM.__new__(USM)
    u.init_key(0xCF222F1FE0748978)
    for seed, size in ((1, 0x140 + 4096), (2, 0x140 + 40), (3, 0x140 + 8), (4, 0x140), (5, 100), (6, 0)):   # the reference needs whole 8-byte words
        data = np.random.default_rng(seed).integers(0, 256, size, dtype=np.uint8).tobytes()
        got = bytes(u.AudioMask(bytearray(data)))
        out["masked"].append({"seed": seed, "size": size, "sha": hashlib.sha256(got).hexdigest()[:16]})
    json.dump(out, open(os.path.join(ROOT, "tests", "golden", "usm_audio.json"), "w"), indent=1)
    print(json.dumps(out, indent=1)[:1200])


if __name__ == "__main__":
    main()
