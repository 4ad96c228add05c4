"""Deterministic synthetic 48 kHz PCM16 corpus.

Pure integer arithmetic, seed = stream id, so every box regenerates the same
bytes without an RNG library: a triangle wave (period/amplitude depend on the
stream and channel) plus LCG noise, with a 50 ms fade-in so the first ADX
block's scale stays below 256 (the reference ADX decoder compares 7 bytes of
"(c)CRI\\0", CriCodecs/adx.cpp:345-348, and the NUL overlaps block 0's scale).

The per-sample LCG  st <- st*1664525 + 1013904223 (mod 2^32)  is evaluated in
closed form  st_n = A_n*st_0 + c*S_n  (A_n = a^n, S_n = sum_{k<n} a^k) so a
whole stream is one vectorised numpy expression.
"""
from __future__ import annotations

import struct
from functools import lru_cache

import numpy as np

SAMPLE_RATE = 48000
DEFAULT_SAMPLES = 96000  # 2 s; multiple of 32 (ADX block) -> 94 HCA frames
_A = 1664525
_C = 1013904223
_M32 = 0xFFFFFFFF


@lru_cache(maxsize=4)
def _lcg_tables(n: int):
    """A_k = a^k and c*S_k for k = 1..n as uint64 arrays (values < 2^32)."""
    a_pow = np.empty(n + 1, dtype=np.uint64)
    a_pow[0] = 1
    # cumulative product mod 2^32 in blocks to stay inside uint64
    cur = 1
    vals = [1]
    for _ in range(n):
        cur = (cur * _A) & _M32
        vals.append(cur)
    a_pow[:] = np.array(vals, dtype=np.uint64)
    s = np.cumsum(a_pow[:-1], dtype=np.uint64) & np.uint64(_M32)  # S_1..S_n
    cs = (s * np.uint64(_C)) & np.uint64(_M32)
    return a_pow[1:].copy(), cs


def lcg_tables_u32(n: int = DEFAULT_SAMPLES):
    """(A_k, c*S_k) for k = 1..n as uint32 arrays (for the device generator)."""
    a, cs = _lcg_tables(n)
    return a.astype(np.uint32), cs.astype(np.uint32)


def channel_samples(sid: int, ch: int, n: int = DEFAULT_SAMPLES) -> np.ndarray:
    """int16[n] samples of channel `ch` of stream `sid`."""
    a_pow, cs = _lcg_tables(n)
    st0 = (0x9E3779B9 ^ ((sid * 2654435761) & _M32) ^ ((ch * 0x85EBCA6B) & _M32)) & _M32
    st = (a_pow * np.uint64(st0) + cs) & np.uint64(_M32)
    noise = ((((st >> np.uint64(16)) & np.uint64(0xFFFF)).astype(np.int64)) - 32768) >> 4
    period = 97 + 13 * ((sid + 3 * ch) % 31)
    amp = 6000 + 500 * ((sid + ch) % 8)
    idx = np.arange(n, dtype=np.int64)
    ph = idx % period
    h = period // 2
    up = (amp * (2 * ph - h)) // h
    down = (amp * (2 * (period - ph) - (period - h))) // (period - h)
    tri = np.where(ph < h, up, down)
    env = np.minimum(4096, (idx * 4096) // 2400)
    x = (env * (tri + noise)) >> 12
    return np.clip(x, -32768, 32767).astype(np.int16)


def pcm(sid: int, channels: int = 2, n: int = DEFAULT_SAMPLES) -> np.ndarray:
    """int16[n, channels] interleaved PCM of stream `sid` (mono = channel 0)."""
    out = np.empty((n, channels), dtype=np.int16)
    for c in range(channels):
        out[:, c] = channel_samples(sid, c, n)
    return out


def wav_header(channels: int, n: int, rate: int = SAMPLE_RATE) -> bytes:
    """Canonical 44-byte RIFF/fmt /data header for 16-bit PCM."""
    data = n * channels * 2
    return struct.pack("<4sI4s4sIHHIIHH4sI", b"RIFF", 36 + data, b"WAVE", b"fmt ", 16, 1, channels,
                       rate, rate * channels * 2, channels * 2, 16, b"data", data)


def wav(sid: int, channels: int = 2, n: int = DEFAULT_SAMPLES, rate: int = SAMPLE_RATE) -> bytes:
    return wav_header(channels, n, rate) + pcm(sid, channels, n).tobytes()


def pcm_batch_torch(sids, channels: int = 2, n: int = DEFAULT_SAMPLES, device="cpu"):
    """int16 tensor [len(sids), n, channels]: the same corpus as `pcm`, vectorised over streams with torch
    (used by bench.py to synthesise thousands of streams on the GPU; plumbing, not part of the codec path)."""
    import torch
    a_pow, cs = _lcg_tables(n)
    A = torch.from_numpy(a_pow.astype(np.int64)).to(device)          # [n]
    CS = torch.from_numpy(cs.astype(np.int64)).to(device)
    sid = torch.as_tensor(list(sids), dtype=torch.int64, device=device).view(-1, 1, 1)   # [S,1,1]
    ch = torch.arange(channels, dtype=torch.int64, device=device).view(1, 1, -1)         # [1,1,C]
    m32 = 0xFFFFFFFF
    st0 = (0x9E3779B9 ^ ((sid * 2654435761) & m32) ^ ((ch * 0x85EBCA6B) & m32)) & m32     # [S,1,C]
    a = A.view(1, -1, 1)
    lo = (a & 0xFFFF) * st0                                          # < 2^48
    hi = (((a >> 16) * st0) & 0xFFFF) << 16
    st = (lo + hi + CS.view(1, -1, 1)) & m32                         # [S,n,C]
    noise = (((st >> 16) & 0xFFFF) - 32768) >> 4
    period = 97 + 13 * ((sid + 3 * ch) % 31)
    amp = 6000 + 500 * ((sid + ch) % 8)
    idx = torch.arange(n, dtype=torch.int64, device=device).view(1, -1, 1)
    ph = idx % period
    h = period // 2
    up = torch.div(amp * (2 * ph - h), h, rounding_mode="floor")
    down = torch.div(amp * (2 * (period - ph) - (period - h)), period - h, rounding_mode="floor")
    tri = torch.where(ph < h, up, down)
    env = torch.clamp((idx * 4096) // 2400, max=4096)
    x = (env * (tri + noise)) >> 12
    return torch.clamp(x, -32768, 32767).to(torch.int16)
