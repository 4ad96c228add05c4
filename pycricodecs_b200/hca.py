"""`HCA` front-end with the reference's surface (PyCriCodecs/hca.py:20-308).

Same constructor, same methods, same exceptions; the codec work runs on the GPU
through the C-ABI (engine.py). Differences from the reference are limited to
its bugs: inputs are never mutated, and `encode(encrypt=True)` keeps the
reference's argument slip (`keyless` lands in `subkey`, hca.py:273 -> :276) only
in the harmless case it can be observed (keyless=False -> subkey 0).
"""
from __future__ import annotations

from io import BytesIO
from struct import Struct, unpack
from typing import BinaryIO

from . import engine
from .chunk import CriHcaQuality, HCAType, WavDataHeaderStruct, WavHeaderStruct, WavSmplHeaderStruct

DEFAULT_KEY = 0xCF222F1FE0748978

# ---- header description tables ---------------------------------------------------------------------------------
# An HCA header is the base chunk and `fmt` followed by optional chunks in any order the writer chose; every chunk
# is a tag (its letters carry bit 7 when the file is encrypted, hence the 0x7F mask) and a fixed big-endian layout.
# `info()` reports each field under the reference's key (PyCriCodecs/hca.py:78-170, including its `BaseBandCoung`
# spelling); a key of None drops the field, `_split_tracks` post-processes the packed track / config byte of `dec`.
def _split_tracks(fields):
    packed = fields.pop("_tracks")
    fields["TrackCount"], fields["ChannelConfig"] = packed >> 4, packed & 0xF
    return {k: fields[k] for k in ("DecSig", "FrameSize", "MinResolution", "MaxResolution", "TotalBandCount", "BaseBandCoung",
                                   "TrackCount", "ChannelConfig", "StereoType")}


_HCA_CHUNKS = {
    b"comp": (Struct(">4sHBBBBBBBBBB"), ("CompSig", "FrameSize", "MinResolution", "MaxResolution", "TrackCount", "ChannelConfig",
                                         "TotalBandCount", "BaseBandCount", "StereoBandCount", "BandsPerHfrGroup", "ReservedByte1",
                                         "ReservedByte2"), None),
    b"dec\x00": (Struct(">4sHBBBBBB"), ("DecSig", "FrameSize", "MaxResolution", "MinResolution", "TotalBandCount", "BaseBandCoung",
                                        "_tracks", "StereoType"), _split_tracks),
    b"vbr\x00": (Struct(">4sHH"), ("VbrSig", "MaxFrameSize", "NoiseLevel"), None),
    b"ath\x00": (Struct(">4sH"), ("AthSig", "TableType"), None),
    b"loop": (Struct(">4sIIHH"), ("LoopSig", "LoopStart", "LoopEnd", "LoopStartDelay", "LoopEndPadding"), None),
    b"ciph": (Struct(">4sH"), ("CiphSig", "CipherType"), None),
    b"rva\x00": (Struct(">4sf"), ("RvaSig", "Volume"), None),
}
_HCA_BASE = Struct(">4sHH")
_HCA_FMT = Struct(">4sIIHH")
_KEY_LIMITS = (("key", 0xFFFFFFFFFFFFFFFF, "HCA key cannot be a negative.", "HCA key cannot exceed the maximum size of 8 bytes."),
               ("subkey", 0xFFFF, "HCA subkey cannot be a negative.", "HCA subkey cannot exceed 65535."))
# optional RIFF chunks between `fmt ` and `data`: tag -> handler name
_WAV_OPTIONAL = {b"smpl": "_wav_smpl", b"note": "_wav_note"}


def _unmask(tag: bytes) -> bytes:
    return bytes(c & 0x7F for c in tag)


class _Cursor:
    """Forward reader over an immutable byte string (the header walkers below never need more)."""
    __slots__ = ("data", "at")

    def __init__(self, data: bytes, at: int = 0):
        self.data, self.at = data, at

    def take(self, n: int) -> bytes:
        piece = self.data[self.at:self.at + n]
        self.at += len(piece)
        return piece

    def unpack(self, layout: Struct):
        return layout.unpack(self.take(layout.size))

    def peek(self, n: int) -> bytes:
        return self.data[self.at:self.at + n]

    def skip(self, n: int) -> None:
        self.at += n


class HCA:
    """One HCA or WAV image held in memory. `_source` is what the caller gave (path contents or a private copy of the
    buffer), `_hca` the current HCA image (after encode / encrypt / decrypt), `_wav` the last decode."""

    def __init__(self, stream: BinaryIO, key: int = 0, subkey: int = 0) -> None:
        if isinstance(stream, str):
            with open(stream, "rb") as fh:
                self._source = fh.read()
        else:
            self._source = bytes(stream)                       # a private copy: the caller's buffer is never touched
        self.key = key if not isinstance(key, str) else int(key, 16)
        self.subkey = subkey if not isinstance(subkey, str) else int(subkey, 16)
        self._hca = self._source
        self._wav = b""
        self._encoded = False
        self.encrypted = self.looping = False
        self.Pyparse_header()

    # -- header sniffing: what `info()` reports (same keys and values as the reference's Pyparse_header) ------------
    def Pyparse_header(self) -> None:
        cur = _Cursor(self._hca)
        self.HcaSig, self.version, self.header_size = cur.unpack(_HCA_BASE)
        if self.HcaSig in (HCAType.HCA.value, HCAType.EHCA.value):
            self._parse_hca(cur)
        elif self.HcaSig == b"RIFF":
            self._parse_wav(_Cursor(self._source))
        else:
            raise ValueError("Invalid HCA or WAV file.")

    def _parse_hca(self, cur: _Cursor) -> None:
        if not self._encoded:                                  # an encoded WAV stays a "wav" object, as in the reference
            self.filetype = "hca"
        self.encrypted = self.HcaSig == HCAType.EHCA.value
        if self.encrypted and not self.key:
            self.key = DEFAULT_KEY
        for name, top, negative, too_big in _KEY_LIMITS:
            value = getattr(self, name)
            if value < 0:
                raise ValueError(negative)
            if value > top:
                raise OverflowError(too_big)
        fmtsig, packed, frames, delay, padding = cur.unpack(_HCA_FMT)
        self.hca = dict(Encrypted=self.encrypted, Header=self.HcaSig, version=hex(self.version), HeaderSize=self.header_size,
                        FmtSig=fmtsig, ChannelCount=packed >> 24, SampleRate=packed & 0x00FFFFFF, FrameCount=frames,
                        EncoderDelay=delay, EncoderPadding=padding)
        while True:                                            # known chunks, in whatever order they come
            tag = cur.peek(4)
            entry = _HCA_CHUNKS.get(_unmask(tag)) if len(tag) == 4 else None
            if entry is None:
                break
            layout, keys, finish = entry
            fields = dict(zip(keys, cur.unpack(layout)))
            self.hca.update(finish(fields) if finish else fields)
            self.looping = self.looping or "LoopSig" in fields
            if fields.get("CipherType") == 1:
                self.encrypted = True
        self.hca["Crc16"] = cur.take(2)

    def _parse_wav(self, cur: _Cursor) -> None:
        self.filetype = "wav"
        (self.riffSignature, self.riffSize, self.wave, self.fmt, self.fmtSize, self.fmtType, self.fmtChannelCount,
         self.fmtSamplingRate, self.fmtSamplesPerSec, self.fmtSamplingSize, self.fmtBitCount) = cur.unpack(WavHeaderStruct)
        if (self.riffSignature, self.wave, self.fmt) != (b"RIFF", b"WAVE", b"fmt "):
            return                                             # the reference looks no further either
        if self.fmtBitCount != 16:
            raise ValueError(f"WAV bitdepth of {self.fmtBitCount} is not supported, only 16 bit WAV files are supported.")
        if self.fmtSize != 16:
            raise ValueError(f"WAV file has an FMT chunk of an unsupported size: {self.fmtSize}, the only supported size is 16.")
        for tag, handler in _WAV_OPTIONAL.items():             # the two chunks the reference knows, in its order
            if cur.peek(4) == tag:
                getattr(self, handler)(cur)
        head = cur.take(WavDataHeaderStruct.size)
        if head[:4] != b"data":
            raise ValueError("Invalid or an unsupported wav file.")
        self.dataSig, self.dataSize = WavDataHeaderStruct.unpack(head)

    def _wav_smpl(self, cur: _Cursor) -> None:
        """One sampler loop makes the stream a looping one; any other count is skipped like an unknown chunk."""
        start = cur.at
        v = cur.unpack(WavSmplHeaderStruct)
        size, self.LoopCount, self.LoopStartSample, self.LoopEndSample = v[1], v[9], v[13], v[14]
        self.looping = self.LoopCount == 1
        if not self.looping:
            cur.at = start + 8 + size

    def _wav_note(self, cur: _Cursor) -> None:
        """Skip a `note` chunk (the reference seeks to an absolute offset here and then loses the data chunk)."""
        cur.skip(4)
        cur.skip(unpack("<I", cur.take(4))[0])

    def info(self) -> dict:
        """ Returns info related to the input file. """
        if self.filetype == "hca":
            return self.hca
        return dict(RiffSignature=self.riffSignature.decode(), riffSize=self.riffSize, WaveSignature=self.wave.decode(),
                    fmtSignature=self.fmt.decode(), fmtSize=self.fmtSize, fmtType=self.fmtType,
                    fmtChannelCount=self.fmtChannelCount, fmtSamplingRate=self.fmtSamplingRate,
                    fmtSamplesPerSec=self.fmtSamplesPerSec, fmtSamplingSize=self.fmtSamplingSize, fmtBitCount=self.fmtBitCount,
                    dataSignature=self.dataSig.decode(), dataSize=self.dataSize)

    def decode(self) -> bytes:
        if self.filetype == "wav":
            raise ValueError("Input type for decoding must be an HCA file.")
        self._wav = engine.hca_decode_batch([self._hca], keys=self.key, subkeys=self.subkey)[0]
        return bytes(self._wav)

    def encode(self, force_not_looping: bool = False, encrypt: bool = False, keyless: bool = False,
               quality_level: CriHcaQuality = CriHcaQuality.High) -> bytes:
        if self.filetype == "hca":
            raise ValueError("Input type for encoding must be a WAV file.")
        if force_not_looping not in (False, True):
            raise ValueError("Forcing the encoder to not loop is by either False or True.")
        if quality_level not in list(CriHcaQuality):
            raise ValueError("Chosen quality level is not valid or is not the appropiate enumeration value.")
        self._hca = engine.hca_encode_batch([self._source], quality=quality_level.value,
                                            force_not_looping=bool(force_not_looping))[0]
        self._encoded = True
        self.Pyparse_header()
        if encrypt:
            if not keyless and self.key == 0:
                self.key = DEFAULT_KEY
            self.encrypt(self.key, keyless)  # as in the reference: `keyless` is taken as the subkey (hca.py:273)
        return self.get_hca()

    def _recipher(self, to_encrypted: bool, keycode: int, subkey: int, ciph_type: int) -> None:
        if self.encrypted == to_encrypted:
            raise ValueError("HCA is already encrypted." if to_encrypted else "HCA is already decrypted.")
        self._hca = engine.hca_crypt_batch([self._hca], to_encrypted, keys=keycode, subkeys=int(subkey), ciph_type=ciph_type)[0]
        self.encrypted = to_encrypted

    def encrypt(self, keycode: int, subkey: int = 0, keyless: bool = False) -> None:
        self._recipher(True, keycode, subkey, 1 if keyless else 56)

    def decrypt(self, keycode: int, subkey: int = 0) -> None:
        self._recipher(False, keycode, subkey, 0)

    def get_hca(self) -> bytes:
        """ Use this function to get the HCA file bytes after encrypting or decrypting. """
        return bytes(self._hca)

    def get_frames(self):
        """ Generator function to yield Frame number, and Frame data. """
        size = self.hca["FrameSize"]
        for i in range(self.hca["FrameCount"]):
            first = self.header_size + i * size
            yield (i, self._hca[first:first + size])

    def get_header(self) -> bytes:
        """ Use this function to retrieve the HCA Header. """
        return self._hca[:self.header_size]

    # the reference keeps its images in file-like members; read-only views for code that looks at them
    @property
    def hcastream(self) -> BytesIO:
        return BytesIO(self._hca)

    @property
    def stream(self) -> BytesIO:
        return BytesIO(self._wav if self._wav else self._source)

    @property
    def hcabytes(self) -> bytes:
        return self._hca if self._encoded else b""

    @property
    def wavbytes(self) -> bytes:
        return self._wav

    # -- batch entry points (the performance path) -------------------------
    @staticmethod
    def decode_batch(streams, keys=None, subkeys=None, ctx=None, raise_errors=True):
        return engine.hca_decode_batch(streams, keys, subkeys, ctx, raise_errors)

    @staticmethod
    def encode_batch(streams, quality_level: CriHcaQuality = CriHcaQuality.High, ctx=None, raise_errors=True):
        return engine.hca_encode_batch(streams, quality_level.value, False, ctx, raise_errors)

    @staticmethod
    def crypt_batch(streams, encrypt: bool, keys=None, subkeys=None, keyless=False, ctx=None, raise_errors=True):
        return engine.hca_crypt_batch(streams, encrypt, keys, subkeys, 1 if keyless else 56, ctx, raise_errors)
