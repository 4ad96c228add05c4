"""GPU: the reference's OWN Python front-end (`PyCriCodecs/adx.py`, `hca.py`, installed unmodified into the git-ignored
baseline/_ref by `pip install --target`) runs on top of the drop-in `CriCodecs` module
(pycricodecs_b200/dropin): same calls, same bytes as the same front-end on top of the reference's compiled extension.

Two worker processes run the same script; the only difference is which `CriCodecs` module is first on sys.path. Each
prints the digests of what `ADX.decode`, `HCA(...).encode / decode / encrypt / decrypt / info / get_frames` returned.
(The reference's `AdxEncode` binding clobbers its block size, adx.cpp:526-527, so `ADX.encode` on the drop-in is compared
with the oracle's harness instead; its `HcaEncode` is only reliable as the first call of a process, which is where the
worker puts it.)"""
import hashlib
import json
import os
import subprocess
import sys
import textwrap

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_PKG = os.path.join(ROOT, "baseline", "_ref")
DROPIN = os.path.join(ROOT, "pycricodecs_b200", "dropin")
KEY = 0xCF222F1FE0748978

WORKER = textwrap.dedent("""
    import sys, json, hashlib
    mode, work = sys.argv[1], sys.argv[2]
    sys.path[:0] = ([%r] if mode == "dropin" else []) + [%r]
    import CriCodecs
    from PyCriCodecs import ADX, HCA, CriHcaQuality            # the reference's package, untouched
    import PyCriCodecs
    h = lambda b: hashlib.sha256(bytes(b)).hexdigest()[:16]
    rd = lambda n: open(work + "/" + n, "rb").read()
    out = {"module": CriCodecs.__file__, "package": PyCriCodecs.__file__}
    wav, adx, hca = rd("in.wav"), rd("in.adx"), rd("in.hca")
    KEY = %d
    try:                                                        # first codec call of the process: HcaEncode
        got = HCA(wav).encode(quality_level=CriHcaQuality.High)
        out["hca_encode"] = h(got) if len(got) == len(hca) else "unreliable"    # the reference's binding reads an
    except Exception as e:                                                       # uninitialised struct (hca.cpp:3468)
        out["hca_encode"] = "unreliable"
    out["info"] = {k: (v.hex() if isinstance(v, bytes) else v) for k, v in HCA(hca).info().items()}
    out["hca_decode"] = h(HCA(hca).decode())
    out["frames"] = h(b"".join(f for _, f in HCA(hca).get_frames()))
    x = HCA(hca)
    x.encrypt(KEY, 0x1234)
    out["encrypt"] = h(x.get_hca())
    out["decode_encrypted"] = h(HCA(x.get_hca(), key=KEY, subkey=0x1234).decode())
    y = HCA(x.get_hca(), key=KEY, subkey=0x1234)
    y.decrypt(KEY, 0x1234)
    out["decrypt"] = h(y.get_hca())
    out["adx_decode"] = h(ADX.decode(adx))
    if mode == "dropin":
        out["adx_encode"] = h(ADX.encode(wav))
        out["adx_encode_v3_hp"] = h(ADX.encode(wav, AdxVersion=3, Highpass_Frequency=300))
        out["hca_encode_low_encrypted"] = h(HCA(wav, key=KEY).encode(encrypt=True, quality_level=CriHcaQuality.Low))
    try:
        HCA(wav).decode()
    except ValueError as e:
        out["wav_decode_error"] = str(e)
    try:
        ADX.decode(b"not an adx file at all........................")
    except Exception as e:
        out["adx_error"] = type(e).__name__ + ": " + str(e)
    print(json.dumps(out))
""")


def _run(tmp_path, mode):
    script = tmp_path / "worker.py"
    script.write_text(WORKER % (DROPIN, REF_PKG, KEY))
    p = subprocess.run([sys.executable, str(script), mode, str(tmp_path)], capture_output=True, text=True, timeout=600, cwd=str(tmp_path))
    assert p.returncode == 0, p.stderr[-3000:]
    return json.loads([ln for ln in p.stdout.splitlines() if ln.startswith("{")][-1])


def test_reference_frontend_runs_unmodified_on_the_dropin(tmp_path, port):
    if not os.path.isdir(os.path.join(REF_PKG, "PyCriCodecs")):
        pytest.skip("baseline/_ref (pip install --target of the reference) is not present")
    from pycricodecs_b200 import synth
    wav = synth.wav(5, 2, 4800 * 3)
    (tmp_path / "in.wav").write_bytes(wav)
    (tmp_path / "in.adx").write_bytes(port.adx_encode(wav)[1])
    hca = port.hca_encode(wav, 1)[1]
    (tmp_path / "in.hca").write_bytes(hca)
    ours = _run(tmp_path, "dropin")
    ref = _run(tmp_path, "reference")
    assert ours["module"].endswith(os.path.join("dropin", "CriCodecs.py")) and "baseline" in ref["module"]
    assert ours["package"] == ref["package"] and "baseline" in ours["package"]        # the same, unmodified, front-end
    h = lambda b: hashlib.sha256(bytes(b)).hexdigest()[:16]
    assert ours["hca_encode"] == h(hca)
    if ref["hca_encode"] != "unreliable":
        assert ref["hca_encode"] == ours["hca_encode"]
    for k in ("info", "hca_decode", "frames", "encrypt", "decode_encrypted", "decrypt", "adx_decode", "wav_decode_error",
              "adx_error"):
        assert ours[k] == ref[k], k
    assert ours["decrypt"] == ours["hca_encode"] and ours["decode_encrypted"] == ours["hca_decode"]
    assert ours["adx_encode"] == h(port.adx_encode(wav)[1])
    assert ours["adx_encode_v3_hp"] == h(port.adx_encode(wav, version=3, highpass=300)[1])
    low = port.hca_encode(wav, 3)[1]
    assert ours["hca_encode_low_encrypted"] == h(port.hca_crypt(low, 1, 56, KEY)[1])
