"""The enums / struct layouts of PyCriCodecs/chunk.py that the HCA front-end uses
(chunk.py:9-12, 42-44, 68-73). Container layouts (UTF/USM/CPK/AWB) are out of scope."""
from enum import Enum
from struct import Struct

WavHeaderStruct = Struct("<4sI4s4sIHHIIHH")
WavSmplHeaderStruct = Struct("<4sIIIIIIIIIIIIIIII")
WavNoteHeaderStruct = Struct("<4sII")
WavDataHeaderStruct = Struct("<4sI")


class HCAType(Enum):
    HCA = b"HCA\x00"            # Header.
    EHCA = b"\xC8\xC3\xC1\x00"  # Encrypted HCA header.


class CriHcaQuality(Enum):
    Highest = 0
    High = 1
    Middle = 2
    Low = 3
    Lowest = 5  # the C++ enum has Lowest = 4, so 5 silently encodes as High (hca.cpp:78, 2211-2227)
