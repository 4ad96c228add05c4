"""USM audio path: key schedule / AudioMask against the reference's USM class
(tools/make_golden_usm.py -> tests/golden/usm_audio.json), batch `load_audio` against per-track oracle encodes."""
import hashlib
import json
import os

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
D = json.load(open(os.path.join(GOLD, "usm_audio.json")))
KEY = 0xCF222F1FE0748978
h = lambda b: hashlib.sha256(bytes(b)).hexdigest()[:16]


def test_key_schedule_matches_the_reference():
    from pycricodecs_b200 import usm_audio
    for e in D["keys"]:
        t = usm_audio.video_mask(e["key"])
        assert t.hex() == e["videomask1"], e["key"]
        assert bytes(v ^ 0xFF for v in t).hex() == e["videomask2"]
        assert usm_audio.audio_mask(e["key"]).hex() == e["audiomask"]
    with pytest.raises(ValueError):
        usm_audio.video_mask("0" * 17)
    with pytest.raises(ValueError):
        usm_audio.video_mask(1.5)


def test_audio_mask_matches_the_reference_and_is_an_involution():
    from pycricodecs_b200 import usm_audio
    mask = usm_audio.audio_mask(KEY)
    for e in D["masked"]:
        data = np.random.default_rng(e["seed"]).integers(0, 256, e["size"], dtype=np.uint8).tobytes()
        got = usm_audio.AudioMask(data, mask)
        assert h(got) == e["sha"], e
        assert got[:0x140] == data[:0x140] and usm_audio.AudioMask(got, mask) == data
    odd = bytes(range(256)) * 2 + b"xyz"                    # a trailing partial word stays as it is
    got = usm_audio.AudioMask(odd, mask)
    assert got[-3:] == b"xyz" and len(got) == len(odd) and usm_audio.AudioMask(got, mask) == odd


def test_track_names():
    from pycricodecs_b200 import usm_audio
    assert usm_audio.track_filenames(["a.wav", b"x", "b.wav", b"y"]) == ["a.wav", "00.sfa", "b.wav", "01.sfa"]
    assert usm_audio.track_filenames(b"x") == ["00.sfa"] and usm_audio.track_filenames("t.wav") == ["t.wav"]
    with pytest.raises(ValueError, match="only HCA and ADX"):
        usm_audio.load_audio([], audio_codec="mp3")


@pytest.mark.gpu
def test_load_audio_encodes_all_tracks_in_one_batch(ctx, port, tmp_path):
    from pycricodecs_b200 import synth, usm_audio
    wavs = [synth.wav(40 + i, 1 + (i & 1), 3000 + 700 * i) for i in range(5)]
    path = tmp_path / "t0.wav"
    path.write_bytes(wavs[0])
    tracks = [str(path)] + wavs[1:]
    launches = ctx.launches
    names, adx = usm_audio.load_audio(tracks, "adx", ctx=ctx)
    assert ctx.launches - launches <= 3
    assert names == [str(path), "00.sfa", "01.sfa", "02.sfa", "03.sfa"]
    for w, a in zip(wavs, adx):
        r, want = port.adx_encode(w, 4, 18, 3, 500, 0, 4)
        assert r == 0 and a == want
    sizes, intervals = usm_audio.sfa_chunk_sizes(adx, "adx")
    assert sizes == [int(48000 // 29.97 // 32) * 18 * (1 + (i & 1)) for i in range(5)] and intervals == [99.9] * 5

    names, hca = usm_audio.load_audio(tracks, "hca", ctx=ctx)
    for w, x in zip(wavs, hca):
        r, want = port.hca_encode(w, 1)
        assert r == 0 and x == want
    sizes, intervals = usm_audio.sfa_chunk_sizes(hca, "hca")
    assert sizes == [int.from_bytes(x[0x1C:0x1E], "big") for x in hca] and intervals == [64] * 5
    # encrypted with the builder's key (type 56); tracks that already are HCA pass through untouched
    _, enc = usm_audio.load_audio(wavs[:2] + [hca[2]], "hca", key=0x1234, encryptAudio=True, ctx=ctx)
    for x, e in zip(hca[:2], enc[:2]):
        assert e == port.hca_crypt(x, 1, 56, 0x1234)[1]
    assert enc[2] == hca[2]
    _, enc0 = usm_audio.load_audio(wavs[:1], "hca", key=0, encryptAudio=True, ctx=ctx)     # key 0 -> the default key
    assert enc0[0] == port.hca_crypt(hca[0], 1, 56, KEY)[1]
