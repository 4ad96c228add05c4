"""Deterministic WAV images in the sample encodings the reference's PCM loader accepts (pcm.cpp:291-327), built from
the synthetic corpus (pure integer generator) so that fixtures and tests agree without shipping the files."""
import struct

import numpy as np

from pycricodecs_b200 import synth


def _riff(fmt_chunk: bytes, data: bytes) -> bytes:
    body = b"WAVE" + b"fmt " + struct.pack("<I", len(fmt_chunk)) + fmt_chunk + b"data" + struct.pack("<I", len(data)) + data
    return b"RIFF" + struct.pack("<I", len(body)) + body


def wav_as(kind: str, sid: int, channels: int, n: int, rate: int = 48000) -> bytes:
    x = synth.pcm(sid, channels, n).astype(np.int64)            # [n][channels] int16 values
    if kind == "u8":
        data, tag, bits, size = ((x >> 8) + 128).astype(np.uint8).tobytes(), 1, 8, 1
    elif kind == "s24":
        v = (x * 256 + (x & 0xFF)).astype(np.int32)
        b = v.astype("<i4").view(np.uint8).reshape(-1, 4)[:, :3]
        data, tag, bits, size = b.tobytes(), 1, 24, 3
    elif kind == "s20in3":                                       # 20 valid bits in a 3-byte container
        v = (x * 16 + (x & 0xF)).astype(np.int32)
        b = v.astype("<i4").view(np.uint8).reshape(-1, 4)[:, :3]
        data, tag, bits, size = b.tobytes(), 1, 20, 3
    elif kind == "s32":
        data, tag, bits, size = (x * 65536 + (x & 0xFFFF) * 3).astype("<i4").tobytes(), 1, 32, 4
    elif kind == "f32":
        f = (x.astype(np.float64) / 29000.0).astype("<f4")       # reaches beyond +-1.0: the clamp is exercised
        f.reshape(-1)[5::997] = np.float32(7e9)                  # and the out-of-range cast
        f.reshape(-1)[7::991] = np.float32(-3e10)
        data, tag, bits, size = f.tobytes(), 3, 32, 4
    elif kind == "f64":
        f = (x.astype(np.float64) / 31000.0).astype("<f8")
        f.reshape(-1)[3::1009] = 1e12
        data, tag, bits, size = f.tobytes(), 3, 64, 8
    elif kind == "ext24":                                        # WAVE_FORMAT_EXTENSIBLE, 24 valid bits, PCM sub-format
        v = (x * 256 + 17).astype(np.int32)
        b = v.astype("<i4").view(np.uint8).reshape(-1, 4)[:, :3]
        data = b.tobytes()
        fmt = struct.pack("<HHIIHHHHI", 0xFFFE, channels, rate, rate * channels * 3, channels * 3, 24, 22, 24, 3)
        fmt += struct.pack("<IHH8s", 1, 0, 0x10, bytes([0x80, 0, 0, 0xAA, 0, 0x38, 0x9B, 0x71]))
        return _riff(fmt, data)
    else:
        raise ValueError(kind)
    fmt = struct.pack("<HHIIHH", tag, channels, rate, rate * channels * size, channels * size, bits)
    return _riff(fmt, data)


KINDS = ("u8", "s24", "s20in3", "s32", "f32", "f64", "ext24")
CASES = [(k, 40 + i, 1 + i % 2, 32 * 40 + 1024 * (i % 3)) for i, k in enumerate(KINDS)]


def loop_wav(sid: int, channels: int, n: int, loop_start: int, loop_end: int, rate: int = 48000) -> bytes:
    """16-bit WAV with a one-loop `smpl` chunk in front of the data (what PCM::GetWaveBuffer itself writes, pcm.cpp:258-265)."""
    pcm = synth.pcm(sid, channels, n).tobytes()
    fmt = struct.pack("<HHIIHH", 1, channels, rate, rate * channels * 2, channels * 2, 16)
    smpl = struct.pack("<9I", 0, 0, 0, 60, 0, 0, 0, 1, 0) + struct.pack("<6I", 0, 0, loop_start, loop_end, 0, 0)
    body = (b"WAVE" + b"fmt " + struct.pack("<I", 16) + fmt + b"smpl" + struct.pack("<I", len(smpl)) + smpl +
            b"data" + struct.pack("<I", len(pcm)) + pcm)
    return b"RIFF" + struct.pack("<I", len(body)) + body


# (sid, channels, samples, loop start, loop end, ADX version)
LOOP_CASES = [(50, 1, 32 * 100, 1000, 3000, 4), (51, 2, 32 * 100, 1000, 3000, 4), (52, 2, 32 * 90 + 7, 33, 2800, 3),
              (53, 2, 32 * 64, 0, 2047, 5)]

# looping HCA encode: (sid, channels, samples, loop start, loop end, quality)
HCA_LOOP_CASES = [(60, 2, 1024 * 12, 3000, 11000, 1), (61, 1, 1024 * 9 + 500, 100, 9000, 1), (62, 2, 1024 * 20, 1024, 1024 * 19, 3),
                  (63, 2, 5000, 0, 4500, 0), (64, 2, 1024 * 40, 17000, 40000, 2)]
