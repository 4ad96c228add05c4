#!/usr/bin/env python3
"""Generate tests/golden/ from the COMPILED REFERENCE (oracle/_ref, built from /root/reference).

Runs only in the dev container (needs oracle/_ref). Writes
  tests/golden/digests.json    sha256 (first 16 hex) of reference outputs on the synthetic corpus
  tests/golden/*.bin           a few small reference-produced streams used as byte-exact vectors
The committed files are what the CPU test-suite and the GPU box check against.
"""
import hashlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
from pycricodecs_b200 import synth  # noqa: E402

KEY = 0xCF222F1FE0748978
h = lambda b: hashlib.sha256(b).hexdigest()[:16]


def main():
    R = oracle.ref()
    out = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out, exist_ok=True)
    d = {"corpus": {}, "kat": {}}
    for sid in (0, 1):
        w2, w1 = synth.wav(sid, 2), synth.wav(sid, 1)
        e = {"wav_stereo": h(w2), "wav_mono": h(w1)}
        for name, w in (("mono", w1), ("stereo", w2)):
            r, a = R.adx_encode(w)
            assert r == 0
            e[f"adx_{name}"] = h(a)
            e[f"adx_{name}_len"] = len(a)
            e[f"adx_{name}_decoded"] = h(R.adx_decode(a))
        for q, qn in enumerate(("highest", "high", "middle", "low")):
            r, x = R.hca_encode(w2, q)
            assert r == 0
            e[f"hca_{qn}"] = h(x)
            e[f"hca_{qn}_len"] = len(x)
            e[f"hca_{qn}_decoded"] = h(R.hca_decode(x))
            r, xm = R.hca_encode(w1, q)
            e[f"hca_mono_{qn}"] = h(xm)
            e[f"hca_mono_{qn}_decoded"] = h(R.hca_decode(xm))
        r, x = R.hca_encode(w2, 1)
        enc = R.hca_crypt(x, 1, 56, KEY)
        e["hca_high_encrypted"] = h(enc)
        e["hca_high_encrypted_subkey_1234"] = h(R.hca_crypt(x, 1, 56, KEY, 0x1234))
        e["hca_high_keyless_type1"] = h(R.hca_crypt(x, 1, 1, 0))
        assert R.hca_crypt(enc, 0, 0, KEY) == x and R.hca_decode(enc, KEY) == R.hca_decode(x)
        d["corpus"][str(sid)] = e
    d["kat"]["cipher56_default_key"] = R.cipher_table(56, KEY).hex()
    d["kat"]["cipher1"] = R.cipher_table(1, 0).hex()
    d["kat"]["crc16_ffff1234"] = R.crc16(b"\xff\xff\x12\x34")
    d["kat"]["adx_coef_48000_500"] = list(R.adx_coefficients(500, 48000))
    d["kat"]["adx_coef_44100_500"] = list(R.adx_coefficients(500, 44100))
    d["kat"]["stereo_samples_2400_2403"] = synth.pcm(0, 2)[2400:2404].tolist()
    json.dump(d, open(os.path.join(out, "digests.json"), "w"), indent=1, sort_keys=True)
    # small byte-exact vectors: 0.1 s streams
    n = 4800
    w = synth.wav(5, 2, n)
    open(os.path.join(out, "s5_stereo_4800.wav"), "wb").write(w)
    open(os.path.join(out, "s5_stereo_4800.adx"), "wb").write(R.adx_encode(w)[1])
    for q, qn in ((1, "high"), (3, "low")):
        x = R.hca_encode(w, q)[1]
        open(os.path.join(out, f"s5_stereo_4800_{qn}.hca"), "wb").write(x)
        open(os.path.join(out, f"s5_stereo_4800_{qn}.hca.wav"), "wb").write(R.hca_decode(x))
    x = R.hca_encode(w, 1)[1]
    open(os.path.join(out, "s5_stereo_4800_high_enc.hca"), "wb").write(R.hca_crypt(x, 1, 56, KEY))
    open(os.path.join(out, "s5_stereo_4800.adx.wav"), "wb").write(R.adx_decode(R.adx_encode(w)[1]))
    print("golden written:", sorted(os.listdir(out)))


if __name__ == "__main__":
    main()
