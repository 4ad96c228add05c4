"""`HCA` front-end with the reference's surface (PyCriCodecs/hca.py:20-308).

Same constructor, same methods, same exceptions; the codec work runs on the GPU
through the C-ABI (engine.py). Differences from the reference are limited to
its bugs: inputs are never mutated, and `encode(encrypt=True)` keeps the
reference's argument slip (`keyless` lands in `subkey`, hca.py:273 -> :276) only
in the harmless case it can be observed (keyless=False -> subkey 0).
"""
from __future__ import annotations

from io import BytesIO, FileIO
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


class HCA:
    def __init__(self, stream: BinaryIO, key: int = 0, subkey: int = 0) -> None:
        if type(stream) == str:
            self.stream = FileIO(stream)
            self.hcastream = FileIO(stream)
        else:
            stream = bytearray(stream).copy()
            self.stream = BytesIO(stream)
            self.hcastream = BytesIO(stream)
        self.key = int(key, 16) if type(key) == str else key
        self.subkey = int(subkey, 16) if type(subkey) == str else subkey
        self.hcabytes = b""
        self.wavbytes = b""
        self.encrypted = False
        self.looping = False
        self.Pyparse_header()

    # -- header sniffing: what `info()` reports (same keys and values as the reference's Pyparse_header) ------------
    def Pyparse_header(self) -> None:
        self.HcaSig, self.version, self.header_size = _HCA_BASE.unpack(self.hcastream.read(_HCA_BASE.size))
        if self.HcaSig in (HCAType.HCA.value, HCAType.EHCA.value):
            self._parse_hca()
        elif self.HcaSig == b"RIFF":
            self._parse_wav()
        else:
            raise ValueError("Invalid HCA or WAV file.")
        self.stream.seek(0)
        self.hcastream.seek(0)

    def _parse_hca(self) -> None:
        if not self.hcabytes:
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
        fmtsig, packed, frames, delay, padding = _HCA_FMT.unpack(self.hcastream.read(_HCA_FMT.size))
        self.hca = dict(Encrypted=self.encrypted, Header=self.HcaSig, version=hex(self.version), HeaderSize=self.header_size,
                        FmtSig=fmtsig, ChannelCount=packed >> 24, SampleRate=packed & 0x00FFFFFF, FrameCount=frames,
                        EncoderDelay=delay, EncoderPadding=padding)
        while True:                                            # known chunks, in whatever order they come
            tag = self.hcastream.read(4)
            entry = _HCA_CHUNKS.get(_unmask(tag)) if len(tag) == 4 else None
            if entry is None:
                self.hcastream.seek(-len(tag), 1)
                break
            layout, keys, finish = entry
            fields = dict(zip(keys, layout.unpack(tag + self.hcastream.read(layout.size - 4))))
            self.hca.update(finish(fields) if finish else fields)
            self.looping = self.looping or "LoopSig" in fields
            if fields.get("CipherType") == 1:
                self.encrypted = True
        self.hca["Crc16"] = self.hcastream.read(2)

    def _parse_wav(self) -> None:
        self.filetype = "wav"
        (self.riffSignature, self.riffSize, self.wave, self.fmt, self.fmtSize, self.fmtType, self.fmtChannelCount,
         self.fmtSamplingRate, self.fmtSamplesPerSec, self.fmtSamplingSize, self.fmtBitCount) = WavHeaderStruct.unpack(
            self.stream.read(WavHeaderStruct.size))
        if (self.riffSignature, self.wave, self.fmt) != (b"RIFF", b"WAVE", b"fmt "):
            return                                             # the reference looks no further either
        if self.fmtBitCount != 16:
            raise ValueError(f"WAV bitdepth of {self.fmtBitCount} is not supported, only 16 bit WAV files are supported.")
        if self.fmtSize != 16:
            raise ValueError(f"WAV file has an FMT chunk of an unsupported size: {self.fmtSize}, the only supported size is 16.")
        for tag, handler in _WAV_OPTIONAL.items():             # the two chunks the reference knows, in its order
            if self.stream.read(4) == tag:
                getattr(self, handler)()
            else:
                self.stream.seek(-4, 1)
        head = self.stream.read(WavDataHeaderStruct.size)
        if head[:4] != b"data":
            raise ValueError("Invalid or an unsupported wav file.")
        self.dataSig, self.dataSize = WavDataHeaderStruct.unpack(head)

    def _wav_smpl(self) -> None:
        """One sampler loop makes the stream a looping one; any other count is skipped like an unknown chunk."""
        self.stream.seek(-4, 1)
        v = WavSmplHeaderStruct.unpack(self.stream.read(WavSmplHeaderStruct.size))
        size, self.LoopCount, self.LoopStartSample, self.LoopEndSample = v[1], v[9], v[13], v[14]
        self.looping = self.LoopCount == 1
        if not self.looping:
            self.stream.seek(8 + size - WavSmplHeaderStruct.size, 1)

    def _wav_note(self) -> None:
        """Skip a `note` chunk (the reference seeks to an absolute offset here and then loses the data chunk)."""
        size = unpack("<I", self.stream.read(4))[0]
        self.stream.seek(size, 1)

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
        self.hcastream.seek(0)
        self.wavbytes = engine.hca_decode_batch([self.hcastream.read()], keys=self.key, subkeys=self.subkey)[0]
        self.stream = BytesIO(self.wavbytes)
        self.hcastream.seek(0)
        return bytes(self.wavbytes)

    def encode(self, force_not_looping: bool = False, encrypt: bool = False, keyless: bool = False,
               quality_level: CriHcaQuality = CriHcaQuality.High) -> bytes:
        if self.filetype == "hca":
            raise ValueError("Input type for encoding must be a WAV file.")
        if force_not_looping not in (False, True):
            raise ValueError("Forcing the encoder to not loop is by either False or True.")
        if quality_level not in list(CriHcaQuality):
            raise ValueError("Chosen quality level is not valid or is not the appropiate enumeration value.")
        self.stream.seek(0)
        self.hcabytes = engine.hca_encode_batch([self.stream.read()], quality=quality_level.value,
                                                force_not_looping=bool(force_not_looping))[0]
        self.hcastream = BytesIO(self.hcabytes)
        self.Pyparse_header()
        if encrypt:
            if self.key == 0 and not keyless:
                self.key = DEFAULT_KEY
            self.encrypt(self.key, keyless)  # as in the reference: `keyless` is taken as the subkey (hca.py:273)
        return self.get_hca()

    def encrypt(self, keycode: int, subkey: int = 0, keyless: bool = False) -> None:
        if self.encrypted:
            raise ValueError("HCA is already encrypted.")
        self.encrypted = True
        enc = engine.hca_crypt_batch([self.get_hca()], True, keys=keycode, subkeys=int(subkey), ciph_type=1 if keyless else 56)[0]
        self.hcastream = BytesIO(enc)

    def decrypt(self, keycode: int, subkey: int = 0) -> None:
        if not self.encrypted:
            raise ValueError("HCA is already decrypted.")
        self.encrypted = False
        dec = engine.hca_crypt_batch([self.get_hca()], False, keys=keycode, subkeys=int(subkey), ciph_type=0)[0]
        self.hcastream = BytesIO(dec)

    def get_hca(self) -> bytes:
        """ Use this function to get the HCA file bytes after encrypting or decrypting. """
        self.hcastream.seek(0)
        fl = self.hcastream.read()
        self.hcastream.seek(0)
        return fl

    def get_frames(self):
        """ Generator function to yield Frame number, and Frame data. """
        self.hcastream.seek(self.header_size, 0)
        for i in range(self.hca["FrameCount"]):
            yield (i, self.hcastream.read(self.hca["FrameSize"]))

    def get_header(self) -> bytes:
        """ Use this function to retrieve the HCA Header. """
        self.hcastream.seek(0)
        header = self.hcastream.read(self.header_size)
        self.hcastream.seek(0)
        return header

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
