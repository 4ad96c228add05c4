"""`ADX` front-end with the reference's surface (PyCriCodecs/adx.py:3-14).

`ADX.decode(data)` and `ADX.encode(data, ...)` are called on the class, exactly
like the reference's; both run the sm_100a kernels through the C-ABI. The
`*_batch` variants take many streams per call (the performance path).
"""
from __future__ import annotations

from . import engine


class ADX:
    """ADX Module for decoding and encoding ADX files, pass the either `adx file` or `wav file` in bytes to either `decode` or `encode` respectively."""

    # Decodes ADX to WAV (CriCodecs.AdxDecode, adx.cpp:546-558).
    def decode(data: bytes) -> bytes:
        """ Decodes ADX to WAV. """
        return engine.adx_decode_batch([bytes(data)])[0]

    # Encodes WAV to ADX (CriCodecs.AdxEncode, adx.cpp:517-544).
    def encode(data: bytes, BitDepth=0x4, Blocksize=0x12, Encoding=3, AdxVersion=0x4, Highpass_Frequency=0x1F4, Filter=0,
               force_not_looping=False) -> bytes:
        """ Encodes WAV to ADX. """
        return engine.adx_encode_batch([bytes(data)], BitDepth=BitDepth, Blocksize=Blocksize, Encoding=Encoding,
                                       Highpass_Frequency=Highpass_Frequency, Filter=Filter, AdxVersion=AdxVersion,
                                       force_not_looping=force_not_looping)[0]

    def decode_batch(streams, ctx=None, raise_errors=True):
        """Decode many ADX streams in one launch; returns a list of WAV images."""
        return engine.adx_decode_batch(streams, ctx, raise_errors)

    def encode_batch(streams, ctx=None, raise_errors=True, **params):
        """Encode many WAV images in one launch; keyword arguments as `encode`."""
        return engine.adx_encode_batch(streams, ctx, raise_errors, **params)
