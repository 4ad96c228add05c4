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

HcaHeaderStruct = Struct(">4sHH")
HcaFmtHeaderStruct = Struct(">4sIIHH")
HcaCompHeaderStruct = Struct(">4sHBBBBBBBBBB")
HcaDecHeaderStruct = Struct(">4sHBBBBBB")
HcaLoopHeaderStruct = Struct(">4sIIHH")
HcaAthHeaderStruct = Struct(">4sH")
HcaVbrHeaderStruct = Struct(">4sHH")
HcaCiphHeaderStruct = Struct(">4sH")
HcaRvaHeaderStruct = Struct(">4sf")

DEFAULT_KEY = 0xCF222F1FE0748978


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

    # -- header sniffing, field for field as the reference's Pyparse_header (hca.py:78-236)
    def Pyparse_header(self) -> None:
        self.HcaSig, self.version, self.header_size = HcaHeaderStruct.unpack(self.hcastream.read(HcaHeaderStruct.size))
        if self.HcaSig in (HCAType.HCA.value, HCAType.EHCA.value):
            if not self.hcabytes:
                self.filetype = "hca"
            self.encrypted = self.HcaSig == HCAType.EHCA.value
            if self.HcaSig == HCAType.EHCA.value and not self.key:
                self.key = DEFAULT_KEY
            elif self.key < 0:
                raise ValueError("HCA key cannot be a negative.")
            elif self.key > 0xFFFFFFFFFFFFFFFF:
                raise OverflowError("HCA key cannot exceed the maximum size of 8 bytes.")
            elif self.subkey < 0:
                raise ValueError("HCA subkey cannot be a negative.")
            elif self.subkey > 0xFFFF:
                raise OverflowError("HCA subkey cannot exceed 65535.")
            fmtsig, temp, framecount, delay, padding = HcaFmtHeaderStruct.unpack(self.hcastream.read(HcaFmtHeaderStruct.size))
            self.hca = dict(Encrypted=self.encrypted, Header=self.HcaSig, version=hex(self.version), HeaderSize=self.header_size,
                            FmtSig=fmtsig, ChannelCount=temp >> 24, SampleRate=temp & 0x00FFFFFF, FrameCount=framecount,
                            EncoderDelay=delay, EncoderPadding=padding)
            while True:
                sig = unpack(">I", self.hcastream.read(4))[0]
                self.hcastream.seek(-4, 1)
                sig = int.to_bytes(sig & 0x7F7F7F7F, 4, "big")
                if sig == b"comp":
                    v = HcaCompHeaderStruct.unpack(self.hcastream.read(HcaCompHeaderStruct.size))
                    self.hca.update(dict(zip(("CompSig", "FrameSize", "MinResolution", "MaxResolution", "TrackCount", "ChannelConfig",
                                              "TotalBandCount", "BaseBandCount", "StereoBandCount", "BandsPerHfrGroup",
                                              "ReservedByte1", "ReservedByte2"), v)))
                elif sig == b"ciph":
                    ciphsig, ciphertype = HcaCiphHeaderStruct.unpack(self.hcastream.read(HcaCiphHeaderStruct.size))
                    if ciphertype == 1:
                        self.encrypted = True
                    self.hca.update(dict(CiphSig=ciphsig, CipherType=ciphertype))
                elif sig == b"loop":
                    self.looping = True
                    v = HcaLoopHeaderStruct.unpack(self.hcastream.read(HcaLoopHeaderStruct.size))
                    self.hca.update(dict(zip(("LoopSig", "LoopStart", "LoopEnd", "LoopStartDelay", "LoopEndPadding"), v)))
                elif sig == b"dec\00":
                    decsig, framesize, maxres, minres, total, base, temp, stereotype = HcaDecHeaderStruct.unpack(
                        self.hcastream.read(HcaDecHeaderStruct.size))
                    self.hca.update(dict(DecSig=decsig, FrameSize=framesize, MinResolution=minres, MaxResolution=maxres,
                                         TotalBandCount=total, BaseBandCoung=base, TrackCount=temp >> 4, ChannelConfig=temp & 0xF,
                                         StereoType=stereotype))
                elif sig == b"ath\00":
                    athsig, tabletype = HcaAthHeaderStruct.unpack(self.hcastream.read(HcaAthHeaderStruct.size))
                    self.hca.update(dict(AthSig=athsig, TableType=tabletype))
                elif sig == b"vbr\00":
                    vbrsig, maxframesize, noiselevel = HcaVbrHeaderStruct.unpack(self.hcastream.read(HcaVbrHeaderStruct.size))
                    self.hca.update(dict(VbrSig=vbrsig, MaxFrameSize=maxframesize, NoiseLevel=noiselevel))
                elif sig == b"rva\00":
                    rvasig, volume = HcaRvaHeaderStruct.unpack(self.hcastream.read(HcaRvaHeaderStruct.size))
                    self.hca.update(dict(RvaSig=rvasig, Volume=volume))
                else:
                    break
            self.hca.update(dict(Crc16=self.hcastream.read(2)))
        elif self.HcaSig == b"RIFF":
            self.filetype = "wav"
            (self.riffSignature, self.riffSize, self.wave, self.fmt, self.fmtSize, self.fmtType, self.fmtChannelCount,
             self.fmtSamplingRate, self.fmtSamplesPerSec, self.fmtSamplingSize, self.fmtBitCount) = WavHeaderStruct.unpack(
                self.stream.read(WavHeaderStruct.size))
            if self.riffSignature == b"RIFF" and self.wave == b"WAVE" and self.fmt == b"fmt ":
                if self.fmtBitCount != 16:
                    raise ValueError(f"WAV bitdepth of {self.fmtBitCount} is not supported, only 16 bit WAV files are supported.")
                elif self.fmtSize != 16:
                    raise ValueError(f"WAV file has an FMT chunk of an unsupported size: {self.fmtSize}, the only supported size is 16.")
                if self.stream.read(4) == b"smpl":
                    self.stream.seek(-4, 1)
                    self.looping = True
                    v = WavSmplHeaderStruct.unpack(self.stream.read(WavSmplHeaderStruct.size))
                    smplesize, self.LoopCount, self.LoopStartSample, self.LoopEndSample = v[1], v[9], v[13], v[14]
                    if self.LoopCount != 1:
                        self.looping = False
                        self.stream.seek(-WavSmplHeaderStruct.size, 1)
                        self.stream.seek(8 + smplesize, 1)
                else:
                    self.stream.seek(-4, 1)
                    self.looping = False
                if self.stream.read(4) == b"note":
                    ln = unpack("<I", self.stream.read(4))[0]
                    self.stream.seek(ln + 4)
                else:
                    self.stream.seek(-4, 1)
                if self.stream.read(4) == b"data":
                    self.stream.seek(-4, 1)
                    self.dataSig, self.dataSize = WavDataHeaderStruct.unpack(self.stream.read(WavDataHeaderStruct.size))
                else:
                    raise ValueError("Invalid or an unsupported wav file.")
        else:
            raise ValueError("Invalid HCA or WAV file.")
        self.stream.seek(0)
        self.hcastream.seek(0)

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
