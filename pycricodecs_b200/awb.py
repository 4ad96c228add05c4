"""AWB (AFS2) audio bank reader whose `extract` / `decode_all` hand every contained HCA stream to the batch decoder
in ONE call -- the batch caller of the hot path (the reference decodes the bank one stream at a time:
PyCriCodecs/awb.py:54-81 loops `HCA(i, key=key, subkey=self.subkey).decode()`).

Mirrors the reference reader (PyCriCodecs/awb.py:9-107): same constructor, attributes (`numfiles`, `align`, `subkey`,
`version`, `ids`, `ofs`, `headersize`), `getfiles()`, `getfile_atindex()`, `extract(decode=False, key=0)` with the
reference's output file names. Header layout: `AWBChunkHeader = "<4sBBHIHH"` (chunk.py:7): magic `AFS2`, version,
offset int size, id int size, file count, alignment, subkey; then the ids, then count + 1 offsets, each rounded up
to the alignment (awb.py:44-45). File i is the bytes between aligned offsets i and i + 1, trailing padding included.
Building banks (AWBBuilder) is container work and stays out of scope.
"""
from __future__ import annotations

import struct
from typing import List, Optional

from .chunk import HCAType
from .hca import HCA

_HEADER = struct.Struct("<4sBBHIHH")
_INT = {1: "B", 2: "H", 4: "I", 8: "Q"}


class AWB:
    __slots__ = ["data", "numfiles", "align", "subkey", "version", "ids", "ofs", "filename", "headersize", "id_alignment"]

    def __init__(self, stream) -> None:
        if isinstance(stream, str):
            self.filename = stream
            with open(stream, "rb") as f:
                self.data = f.read()
        else:
            self.filename = ""
            self.data = bytes(stream)
        self.readheader()

    def readheader(self) -> None:
        if len(self.data) < _HEADER.size:
            raise ValueError("Invalid AWB header.")
        magic, self.version, offset_intsize, id_intsize, self.numfiles, self.align, self.subkey = _HEADER.unpack_from(self.data, 0)
        if magic != b"AFS2":
            raise ValueError("Invalid AWB header.")
        if offset_intsize not in _INT or id_intsize not in _INT:
            raise ValueError("Unknown int size.")
        if self.align == 0:
            raise ValueError("Invalid AWB header.")
        pos = _HEADER.size
        need = pos + id_intsize * self.numfiles + offset_intsize * (self.numfiles + 1)
        if need > len(self.data):
            raise ValueError("Invalid AWB header.")
        self.ids = list(struct.unpack_from(f"<{self.numfiles}{_INT[id_intsize]}", self.data, pos))
        pos += id_intsize * self.numfiles
        raw = struct.unpack_from(f"<{self.numfiles + 1}{_INT[offset_intsize]}", self.data, pos)
        a = self.align
        self.ofs = [o if o % a == 0 else o + (a - o % a) for o in raw]
        self.headersize = need if need % a == 0 else need + (a - need % a)
        self.id_alignment = id_intsize

    def getfiles(self) -> List[bytes]:
        """Every file of the bank, in order (the reference yields them from a generator, awb.py:83-88)."""
        return [self.data[self.ofs[i - 1]:self.ofs[i]] for i in range(1, len(self.ofs))]

    def getfile_atindex(self, index: int) -> bytes:
        return self.data[self.ofs[index]:self.ofs[index + 1]]

    @staticmethod
    def _is_hca(blob: bytes) -> bool:
        return blob.startswith(HCAType.HCA.value) or blob.startswith(HCAType.EHCA.value)

    def decode_all(self, key: int = 0, ctx=None, raise_errors: bool = True) -> List[Optional[bytes]]:
        """WAV bytes of every HCA stream of the bank (None for entries that are not HCA), decoded in one batch call
        with the bank's subkey (awb.py:70)."""
        files = self.getfiles()
        idx = [i for i, f in enumerate(files) if self._is_hca(f)]
        out: List[Optional[bytes]] = [None] * len(files)
        if idx:
            got = HCA.decode_batch([files[i] for i in idx], keys=key, subkeys=self.subkey, ctx=ctx, raise_errors=raise_errors)
            for i, g in zip(idx, got):
                out[i] = g
        return out

    def extract(self, decode: bool = False, key: int = 0, ctx=None) -> List[str]:
        """Write the files next to the bank (or into the working directory for in-memory banks) under the reference's
        names (awb.py:54-81); returns the names written."""
        files = self.getfiles()
        wavs = self.decode_all(key, ctx) if decode else None
        stem = self.filename.rsplit(".", 1)[0] + "_" if self.filename else ""
        written = []
        for count, blob in enumerate(files):
            if self._is_hca(blob):
                name, payload = (f"{stem}{count}.wav", wavs[count]) if decode else (f"{stem}{count}.hca", blob)
            else:
                name, payload = f"{stem}{count}.dat", blob          # "Probably ADX." (awb.py:64)
            with open(name, "wb") as f:
                f.write(payload)
            written.append(name)
        return written
