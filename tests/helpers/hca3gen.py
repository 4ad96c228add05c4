"""Synthetic HCA v3.0 streams for the decoder tests.

The reference can only ENCODE v2.0 (hca.cpp:2414), so v3.0 inputs -- extra scalefactors for the derived HFR scales,
delta-coded intensities, resolution 0 bands rebuilt by the noise generator (hca.cpp:1297-1307, 1382-1424, 1602-1635) --
are built here: a valid header, then frames whose side information (noise level, scalefactors, intensities) is written
field by field and whose spectra are seeded random bits. Any bit pattern is a valid run of codes, so the compiled
reference decodes these streams and its output is the expected value.
"""
import numpy as np


def crc16(data: bytes) -> int:
    crc = 0
    for b in data:
        crc ^= b << 8
        for _ in range(8):
            crc = ((crc << 1) ^ 0x8005) & 0xFFFF if crc & 0x8000 else (crc << 1) & 0xFFFF
    return crc


_CRC_TAB = None


def crc16_fast(data: bytes) -> int:
    global _CRC_TAB
    if _CRC_TAB is None:
        _CRC_TAB = [crc16(bytes([i])) for i in range(256)]
    crc = 0
    for b in data:
        crc = ((crc << 8) & 0xFFFF) ^ _CRC_TAB[(crc >> 8) ^ b]
    return crc


class Bits:
    def __init__(self):
        self.v, self.n = 0, 0

    def put(self, value, count):
        assert 0 <= value < (1 << count) or count == 0
        self.v = (self.v << count) | value
        self.n += count

    def to_bytes(self, size, rng):
        """Pad with random bits up to `size` bytes (or cut, if the side information alone is longer)."""
        total = size * 8
        if self.n < total:
            pad = total - self.n
            self.v = (self.v << pad) | int.from_bytes(rng.bytes((pad + 7) // 8), "big") >> ((-pad) % 8)
            self.n = total
        return (self.v >> (self.n - total)).to_bytes(size, "big")


def header(version=0x0300, channels=2, rate=44100, frames=8, delay=128, padding=0, frame_size=1024, min_res=0, max_res=15,
           tracks=1, config=0, total=128, base=40, stereo=30, bands_per_hfr=4, ciph=0, dec=False, ath=None):
    """dec=True writes the v1.x `dec` chunk instead of `comp` (hca.cpp:710-727: band counts minus one, track count and
    channel config in one byte, a stereo type flag; stereo bands = total - base, no HFR groups); ath = 0 / 1 adds an
    `ath` chunk (absent: type 1 below v2.0, else 0, hca.cpp:745-756)."""
    h = b"HCA\x00" + version.to_bytes(2, "big") + b"\x00\x00"
    h += b"fmt\x00" + bytes([channels]) + rate.to_bytes(3, "big") + frames.to_bytes(4, "big") + delay.to_bytes(2, "big") + padding.to_bytes(2, "big")
    if dec:
        assert bands_per_hfr == 0 and total - base == stereo
        h += b"dec\x00" + frame_size.to_bytes(2, "big") + bytes([min_res, max_res, total - 1, base - 1, (tracks << 4) | config, 1 if stereo else 0])
    else:
        h += b"comp" + frame_size.to_bytes(2, "big") + bytes([min_res, max_res, tracks, config, total, base, stereo, bands_per_hfr, 0, 0])
    if ath is not None:
        h += b"ath\x00" + ath.to_bytes(2, "big")
    h += b"ciph" + ciph.to_bytes(2, "big")
    h += b"pad\x00" + bytes(8)
    size = len(h) + 2
    h = h[:6] + size.to_bytes(2, "big") + h[8:]
    return h + crc16_fast(h).to_bytes(2, "big")


def stream(seed, frames=8, channels=2, frame_size=1024, total=128, base=40, stereo=30, bands_per_hfr=4, min_res=0, max_res=15,
           delay=128, version=0x0300, level=(40, 110), rate=44100, sf_max=44, dec=False, ath=None, kept=0.0):
    """One v3.0 stream (or, with version <= 0x0200, a stream in the older bitstream layout; dec / ath: see header()).
    Channel pairs are primary/secondary when stereo > 0 (hca.cpp:909-960, 2 channels per track).
    kept > 0: that share of the secondary-channel frames stop coding their intensities early, so the decoder keeps the
    previous frame's values for the rest -- v <= 2.0: a first index of 15 (hca.cpp:1368-1372); v3.0: a delta that
    leaves 0..15 (hca.cpp:1410-1412, the caller at :1185 carries on). Streams with kept == 0 are unchanged."""
    rng = np.random.default_rng(seed)
    out = header(version, channels, rate, frames, delay, 0, frame_size, min_res, max_res, 1, 0, total, base, stereo, bands_per_hfr, 0, dec, ath)
    # channel roles for one track and channel_config 0 (hca.cpp:909-960): 1 primary, 2 secondary, 0 discrete
    roles = {1: [0], 2: [1, 2], 3: [1, 2, 0], 4: [1, 2, 1, 2]}[channels] if stereo > 0 else [0] * channels
    rest = total - base - stereo
    groups = 0 if bands_per_hfr == 0 else (rest + bands_per_hfr - 1) // bands_per_hfr
    for _ in range(frames):
        b = Bits()
        b.put(0xFFFF, 16)
        b.put(int(rng.integers(level[0], level[1])), 9)
        b.put(int(rng.integers(0, 128)), 7)
        for c in range(channels):
            secondary = roles[c] == 2
            count = base if secondary else base + stereo
            if not secondary and groups > 0 and version > 0x0200:
                count += groups
            mode = int(rng.integers(0, 8))
            if mode == 0:                                   # no scalefactors
                b.put(0, 3)
            elif mode <= 3:                                 # fixed
                b.put(6 + (mode & 1), 3)
                for _i in range(count):
                    b.put(int(rng.integers(0, sf_max)), 6)
            else:                                           # delta coded, with escapes
                db = int(rng.integers(1, 6))
                esc = (1 << db) - 1
                b.put(db, 3)
                v = int(rng.integers(0, sf_max))
                b.put(v, 6)
                for _i in range(1, count):
                    d = int(rng.integers(0, esc + 1))
                    nv = v + d - (esc >> 1)
                    if d == esc or nv < 0 or nv >= sf_max or rng.random() < 0.1:
                        v = int(rng.integers(0, sf_max))
                        b.put(esc, db)
                        b.put(v, 6)
                    else:
                        v = nv
                        b.put(d, db)
            if secondary:
                stop = kept > 0 and rng.random() < kept
                if version <= 0x0200:
                    if stop:
                        b.put(15, 4)                        # peeked, not consumed: the spectra start at these bits
                    else:
                        for _i in range(8):
                            b.put(int(rng.integers(0, 15)), 4)
                elif stop:
                    up = bool(rng.integers(0, 2))
                    db = 2 if up else int(rng.integers(1, 3))   # 1-bit deltas cannot leave the range, 2-bit ones only downwards
                    bmax, bits = (2 << db) - 1, db + 1
                    at = int(rng.integers(1, 8))            # the delta that leaves the range
                    v = 14 if up else 0
                    b.put(v, 4)
                    b.put(db, 2)
                    for _i in range(1, at):
                        b.put(bmax >> 1, bits)              # delta 0
                    b.put(bmax - 1 if up else 0, bits)      # 14 + (bmax >> 1) > 15, or 0 - (bmax >> 1) wraps past 15
                else:
                    kind = int(rng.integers(0, 6))
                    if kind == 0:                           # "15": all intensities 7
                        b.put(15, 4)
                    else:
                        v = int(rng.integers(0, 15))
                        b.put(v, 4)
                        db = 3 if kind == 1 else int(rng.integers(0, 3))
                        b.put(db, 2)
                        if db == 3:
                            for _i in range(7):
                                b.put(int(rng.integers(0, 16)), 4)
                        else:
                            bmax, bits = (2 << db) - 1, db + 1
                            for _i in range(7):
                                d = int(rng.integers(0, bmax + 1))
                                nv = v + d - (bmax >> 1)
                                if d == bmax or nv < 0 or nv > 15:
                                    v = int(rng.integers(0, 16))
                                    b.put(bmax, bits)
                                    b.put(v, 4)
                                else:
                                    v = nv
                                    b.put(d, bits)
            elif version <= 0x0200:
                for _i in range(groups):
                    b.put(int(rng.integers(0, 64)), 6)
        body = b.to_bytes(frame_size - 2, rng)
        out += body + crc16_fast(body).to_bytes(2, "big")
    return out
