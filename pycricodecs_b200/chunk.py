"""The enums / struct layouts of PyCriCodecs/chunk.py that the HCA front-end uses
(chunk.py:9-12, 42-44, 68-73), plus the @UTF header / value types the ACB cue-sheet reader needs (chunk.py:4, 35-37,
54-66). Names and values are the reference's public surface (callers pass these enums); the other container layouts
(USM/CPK) are out of scope."""
from enum import Enum
from struct import Struct


def _layout(order: str, *fields: str) -> Struct:
    """A struct layout from its fields, one format code (with optional repeat count) per field."""
    return Struct(order + "".join(fields))


# RIFF / WAVE pieces (little-endian): tag, size, then the chunk's fixed fields
WavHeaderStruct = _layout("<", "4s", "I", "4s",                # RIFF, size, WAVE
                          "4s", "I", "H", "H", "I", "I", "H", "H")   # fmt : size, type, channels, rate, bytes/s, block align, bits
WavSmplHeaderStruct = _layout("<", "4s", "I", "7I", "I", "I", "6I")  # smpl: size, 7 sampler words, loop count, extra; first loop: id, type, start, end, fraction, play count
WavNoteHeaderStruct = _layout("<", "4s", "I", "I")
WavDataHeaderStruct = _layout("<", "4s", "I")

# @UTF table header (big-endian): magic, table size, rows / strings / data offsets, name, columns, row length, rows
UTFChunkHeader = _layout(">", "4s", "I", "I", "I", "I", "I", "H", "H", "I")

UTFType = Enum("UTFType", {"UTF": b"@UTF", "EUTF": b"\x1F\x9E\xF3\xF5"})                  # plain / masked table magic
UTFTypeValues = Enum("UTFTypeValues", [(name, code) for code, name in enumerate(
    "uchar char ushort short uint int ullong llong float double string bytes".split())])  # column type nibble
HCAType = Enum("HCAType", {"HCA": b"HCA\x00", "EHCA": b"\xC8\xC3\xC1\x00"})               # plain / bit-7-masked signature
# the C++ enum has Lowest = 4, so 5 silently encodes as High (hca.cpp:78, 2211-2227)
CriHcaQuality = Enum("CriHcaQuality", {"Highest": 0, "High": 1, "Middle": 2, "Low": 3, "Lowest": 5})
