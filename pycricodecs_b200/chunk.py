"""The enums / struct layouts of PyCriCodecs/chunk.py that the HCA front-end uses
(chunk.py:9-12, 42-44, 68-73), plus the @UTF header / value types the ACB cue-sheet reader needs (chunk.py:4, 35-37,
54-66). The other container layouts (USM/CPK) are out of scope."""
from enum import Enum
from struct import Struct

WavHeaderStruct = Struct("<4sI4s4sIHHIIHH")
WavSmplHeaderStruct = Struct("<4sIIIIIIIIIIIIIIII")
WavNoteHeaderStruct = Struct("<4sII")
WavDataHeaderStruct = Struct("<4sI")


UTFChunkHeader = Struct(">4sIIIIIHHI")


class UTFType(Enum):
    UTF = b"@UTF"               # Header.
    EUTF = b"\x1F\x9E\xF3\xF5"  # Encrypted @UTF header.


class UTFTypeValues(Enum):
    uchar = 0
    char = 1
    ushort = 2
    short = 3
    uint = 4
    int = 5
    ullong = 6
    llong = 7
    float = 8
    double = 9
    string = 10
    bytes = 11


class HCAType(Enum):
    HCA = b"HCA\x00"            # Header.
    EHCA = b"\xC8\xC3\xC1\x00"  # Encrypted HCA header.


class CriHcaQuality(Enum):
    Highest = 0
    High = 1
    Middle = 2
    Low = 3
    Lowest = 5  # the C++ enum has Lowest = 4, so 5 silently encodes as High (hca.cpp:78, 2211-2227)
