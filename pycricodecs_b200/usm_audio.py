"""Audio side of the USM container: the part of PyCriCodecs/usm.py that feeds the codecs.

`load_audio` is `USMBuilder.load_audio` (usm.py:437-477) with the per-track encode loop replaced by ONE batch call:
WAV tracks become ADX (v4, mode 3, not looping) or HCA (quality High, not looping, optionally encrypted with type 56
and the builder's key -- `HCA.encode(encrypt=True)` falls back to the library's default key when the key is 0,
hca.py:268-273). `sfa_chunk_sizes` is `prepare_SFA` (usm.py:1152-1177). `audio_mask` / `AudioMask` are the key schedule
and the XOR that USM applies to ADX payloads (usm.py:47-118, 313-322); they are container glue on host bytes, not a
codec kernel. Muxing video, the CRID / @SFA chunk headers and the rest of the container stay out of scope.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple, Union

import numpy as np

from . import engine
from .chunk import HCAType
from .hca import DEFAULT_KEY, HCA

Track = Union[str, bytes, bytearray]


def _read(track: Track) -> bytes:
    if isinstance(track, str):
        with open(track, "rb") as f:
            return f.read()
    return bytes(track)


def track_filenames(audio) -> List[str]:
    """Names the builder records per track: the path, or 00.sfa, 01.sfa ... for in-memory tracks (usm.py:438-451)."""
    tracks = audio if isinstance(audio, list) else [audio]
    names, count = [], 0
    for t in tracks:
        if isinstance(t, str):
            names.append(t)
        else:
            names.append("{:02d}.sfa".format(count))
            count += 1
    return names


def load_audio(audio, audio_codec: str = "adx", key: int = 0, encryptAudio: bool = False, ctx=None) -> Tuple[List[str], List[bytes]]:
    """Encoded streams of all tracks, in order: every WAV of the list goes through the encoder in one batch."""
    tracks = audio if isinstance(audio, list) else [audio]
    blobs = [_read(t) for t in tracks]
    if audio_codec == "adx":
        streams = engine.adx_encode_batch(blobs, ctx, AdxVersion=4, Encoding=3, force_not_looping=True) if blobs else []
    elif audio_codec == "hca":
        is_hca = [b[:4] in (HCAType.HCA.value, HCAType.EHCA.value) for b in blobs]
        wavs = [b for b, h in zip(blobs, is_hca) if not h]
        enc = engine.hca_encode_batch(wavs, quality=1, force_not_looping=True, ctx=ctx) if wavs else []
        if encryptAudio and enc:
            enc = engine.hca_crypt_batch(enc, True, keys=key if key else DEFAULT_KEY, subkeys=0, ciph_type=56, ctx=ctx)
        it = iter(enc)
        streams = [b if h else next(it) for b, h in zip(blobs, is_hca)]
    else:
        raise ValueError("Supported audio codecs in USM are only HCA and ADX.")
    return track_filenames(audio), list(streams)


def sfa_chunk_sizes(streams: Sequence[bytes], audio_codec: str = "adx", video_codec: str = "vp9"):
    """(chunk size, base interval) per encoded stream. ADX: the blocks of one 29.97 Hz tick,
    int(rate // 29.97 // 32) * blocksize * channels; HCA: one frame, interval 64."""
    sizes, intervals = [], []
    for s in streams:
        if audio_codec == "adx":
            blocksize, channels, rate = s[5], s[7], int.from_bytes(s[8:12], "big")
            sizes.append(int(rate // 29.97 // 32) * (blocksize * channels))
            intervals.append(99.9 if video_codec == "vp9" else 100)
        else:
            h = HCA(s)
            sizes.append(h.hca["FrameSize"])
            intervals.append(64)
    return sizes, intervals


# (destination, operation, a, b): a / b index the table unless the operation says "k" (constant). Bytes 0..8 come from
# the key: k1 = low 32 bits, k2 = high 32 bits, both big-endian (usm.py:55-69).
_SCHEDULE = (
    (0x09, "sub", 0x01, 0x07), (0x0A, "xork", 0x02, 0xFF), (0x0B, "xork", 0x01, 0xFF), (0x0C, "add", 0x0B, 0x09),
    (0x0D, "sub", 0x08, 0x03), (0x0E, "xork", 0x0D, 0xFF), (0x0F, "sub", 0x0A, 0x0B), (0x10, "sub", 0x08, 0x0F),
    (0x11, "xor", 0x10, 0x07), (0x12, "xork", 0x0F, 0xFF), (0x13, "xork", 0x03, 0x10), (0x14, "addk", 0x04, -0x32),
    (0x15, "addk", 0x05, 0xED), (0x16, "xork", 0x06, 0xF3), (0x17, "sub", 0x13, 0x0F), (0x18, "add", 0x15, 0x07),
    (0x19, "ksub", 0x21, 0x13), (0x1A, "xor", 0x14, 0x17), (0x1B, "add", 0x16, 0x16), (0x1C, "addk", 0x17, 0x44),
    (0x1D, "add", 0x03, 0x04), (0x1E, "sub", 0x05, 0x16), (0x1F, "xor", 0x1D, 0x13),
)


def video_mask(key) -> bytes:
    """The 32-byte table both masks derive from (`videomask1`, usm.py:47-108)."""
    if isinstance(key, str):
        if len(key) > 16:
            raise ValueError("Invalid input key.")
        key = int(key.rjust(16, "0"), 16)
    elif not isinstance(key, int):
        raise ValueError("Invalid key format, must be either a string or an integer.")
    k1 = (key & 0xFFFFFFFF).to_bytes(4, "big")
    k2 = ((key >> 32) & 0xFFFFFFFF).to_bytes(4, "big")
    t = [0] * 32
    t[0:9] = [k1[3], k1[2], k1[1], k1[0] - 0x34, k2[3] + 0xF9, k2[2] ^ 0x13, k2[1] + 0x61, k1[3] ^ 0xFF, k1[1] + k1[2]]
    t = [v & 0xFF for v in t]
    for dst, op, a, b in _SCHEDULE:
        v = {"add": lambda: t[a] + t[b], "sub": lambda: t[a] - t[b], "xor": lambda: t[a] ^ t[b],
             "addk": lambda: t[a] + b, "xork": lambda: t[a] ^ b, "ksub": lambda: a - t[b]}[op]()
        t[dst] = v & 0xFF
    return bytes(t)


def audio_mask(key) -> bytes:
    """usm.py:109-117: odd bytes spell "URUC", even bytes are the complemented table."""
    t = video_mask(key)
    return bytes(b"URUC"[(x >> 1) & 3] if x & 1 else t[x] ^ 0xFF for x in range(32))


def AudioMask(memObj: bytes, mask: bytes) -> bytes:
    """usm.py:313-322: the first 0x140 bytes stay, then every whole 8-byte word is XORed with mask word (index mod 4)."""
    data = bytes(memObj)
    body = np.frombuffer(data, np.uint8, offset=min(0x140, len(data))).copy()
    words = len(body) // 8
    if words:
        m = np.resize(np.frombuffer(bytes(mask), np.uint8), words * 8)
        body[:words * 8] ^= m
    return data[:0x140] + body.tobytes()
