"""Drop-in replacement of the reference's CPython extension module `CriCodecs` (CriCodecs/CriCodecs.cpp:8-28).

Put this directory on sys.path *instead of* the compiled `CriCodecs` extension and the reference's own front-end
(`PyCriCodecs/adx.py`, `PyCriCodecs/hca.py`) runs unmodified on the GPU engine:

    sys.path.insert(0, ".../pycricodecs_b200/dropin")
    from PyCriCodecs import ADX, HCA          # the reference's Python package, untouched

Same callables, same positional arguments, same return type (new `bytes`), same exception types and messages.
Differences are limited to the reference's bugs: arguments are honoured (the reference's AdxEncode clobbers
blocksize, adx.cpp:526-527), inputs are never mutated (HcaCrypt does, hca.cpp:3298-3300), no sticky error state.
CriLaylaDecompress / CriLaylaCompress (CPK archive compression) are outside the accelerated path.
"""
import os
import sys

_root = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if _root not in sys.path:
    sys.path.insert(0, _root)

from pycricodecs_b200 import engine as _engine  # noqa: E402


def _plain_errors(fn):
    """The reference raises plain ValueError / NotImplementedError (py_codec_err, hca.cpp:3252-3268; adx.cpp:32-38): callers that
    compare exception types must see exactly those, not the engine's ValueError subclass."""
    def call(*a, **k):
        try:
            return fn(*a, **k)
        except _engine.CriError as e:
            raise ValueError(str(e)) from None
    call.__doc__, call.__name__ = fn.__doc__, fn.__name__
    return call



@_plain_errors
def AdxDecode(data):
    """METH_O, adx.cpp:546: ADX image -> WAV image."""
    return _engine.adx_decode_batch([bytes(data)])[0]


@_plain_errors
def AdxEncode(data, bitdepth, blocksize, encoding, highpass, filter, adx_ver, force_no_looping):
    """"y#IIIIIIp", adx.cpp:527 (argument order as PyCriCodecs/adx.py:14 passes them)."""
    return _engine.adx_encode_batch([bytes(data)], BitDepth=bitdepth, Blocksize=blocksize, Encoding=encoding,
                                    Highpass_Frequency=highpass, Filter=filter, AdxVersion=adx_ver,
                                    force_not_looping=bool(force_no_looping))[0]


@_plain_errors
def HcaDecode(data, header_size, key, subkey):
    """"y#IKH", hca.cpp:3352. header_size is re-read from the stream (the reference trusts the caller's value)."""
    return _engine.hca_decode_batch([bytes(data)], keys=int(key) & 0xFFFFFFFFFFFFFFFF, subkeys=int(subkey) & 0xFFFF)[0]


@_plain_errors
def HcaEncode(data, force_nolooping, quality):
    """"y*II", hca.cpp:3463."""
    return _engine.hca_encode_batch([bytes(data)], quality=int(quality), force_not_looping=bool(force_nolooping))[0]


@_plain_errors
def HcaCrypt(buf, crypt, header_size, ciph_type, key, subkey):
    """"OIIIKH", hca.cpp:3280: crypt 1 = encrypt with `ciph_type` (56 keyed / 1 keyless), 0 = decrypt."""
    return _engine.hca_crypt_batch([bytes(buf)], bool(crypt), keys=int(key) & 0xFFFFFFFFFFFFFFFF, subkeys=int(subkey) & 0xFFFF,
                                   ciph_type=int(ciph_type) if crypt else 0)[0]


def CriLaylaDecompress(*_a, **_k):
    raise NotImplementedError("CRILAYLA (CPK archive compression) is outside the accelerated ADX/HCA path")


CriLaylaCompress = CriLaylaDecompress
