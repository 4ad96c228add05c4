"""Read-only @UTF table reader: what the ACB cue-sheet front-end (acb.py) needs from PyCriCodecs/utf.py:7-187.

`UTF(stream).get_payload()` returns the reference's payload shape: one dict per row, column name ->
`(UTFTypeValues, value)`. Layout (all integers big-endian, chunk.py:4 `">4sIIIIIHHI"`): magic `@UTF`, table size,
then -- every offset counted from byte 8 -- rows offset, string-pool offset, data offset, table-name string offset,
column count, row length, row count. Each column is one flag byte (high nibble = storage, low nibble = type) and a
name offset; storage 0x1 has no value (the reference reports `None`, `"<NULL>"` for strings, `b""` for bytes:
utf.py:114-127), 0x3 carries one constant after the name, 0x5 has a cell in every row. Strings are offsets into the
NUL-separated pool, byte cells are (offset into the data area, length). The XOR-masked variant (magic 1F 9E F3 F5,
utf.py:36-50) is unmasked first. Building tables (UTFBuilder) is container work and stays out of scope.
"""
from __future__ import annotations

import struct
from typing import List

import numpy as np

from .chunk import UTFChunkHeader, UTFType, UTFTypeValues

_FMT = ("B", "b", "H", "h", "I", "i", "Q", "q", "f", "d", "I", "II")     # utf.py:159-166
_TYPES = list(UTFTypeValues)


def _unmask(data: bytes) -> bytes:
    """utf.py:38-44: byte i is XORed with the low byte of 0x655F * 0x4115**i (mod 2**32). Only the low bytes of the
    seed and the multiplier reach that byte, so the mask is 0x5F * 0x15**i mod 256, which repeats every 64 bytes."""
    period = np.empty(64, np.uint8)
    v = 0x5F
    for i in range(64):
        period[i] = v
        v = (v * 0x15) & 0xFF
    assert v == 0x5F
    mask = np.resize(period, len(data))
    return (np.frombuffer(data, np.uint8) ^ mask).tobytes()


class UTF:
    __slots__ = ["magic", "table_size", "rows_offset", "string_offset", "data_offset", "table_name", "num_columns",
                 "row_length", "num_rows", "data", "table", "_payload", "encoding"]

    def __init__(self, stream) -> None:
        if isinstance(stream, str):
            with open(stream, "rb") as f:
                data = f.read()
        else:
            data = bytes(stream)
        if data[:4] == UTFType.EUTF.value:
            data = _unmask(data)
            if data[:4] != UTFType.UTF.value:
                raise Exception("Decryption error.")
        if len(data) < UTFChunkHeader.size or data[:4] != UTFType.UTF.value:
            raise ValueError("UTF chunk is not present.")
        self.data = data
        (self.magic, self.table_size, self.rows_offset, self.string_offset, self.data_offset, name_ptr, self.num_columns,
         self.row_length, self.num_rows) = UTFChunkHeader.unpack_from(data, 0)
        if max(self.rows_offset, self.string_offset, self.data_offset) + 8 > len(data) or self.string_offset > self.data_offset:
            raise ValueError("UTF chunk is truncated.")
        self.encoding = "utf-8"
        self.table_name = self._string(name_ptr)
        self._read()

    def _string(self, ptr: int) -> str:
        lo = 8 + self.string_offset + ptr
        hi = 8 + self.data_offset
        if lo >= hi:
            raise Exception("Failed string lookup.")
        end = self.data.find(b"\x00", lo, hi)
        raw = self.data[lo:end if end >= 0 else hi]
        for enc in ("utf-8", "shift-jis", "utf-16"):                     # utf.py:96-110
            try:
                text = raw.decode(enc)
                if enc != "utf-8":
                    self.encoding = enc
                return text
            except UnicodeDecodeError:
                continue
        raise ValueError(f"String of unknown encoding: {raw!r}")

    def _value(self, kind: int, at: int):
        """The (typed) value stored at byte `at`, and the bytes it occupies."""
        fmt = ">" + _FMT[kind]
        raw = struct.unpack_from(fmt, self.data, at)
        if kind == 0xA:
            return (UTFTypeValues.string, self._string(raw[0])), 4
        if kind == 0xB:
            lo = 8 + self.data_offset + raw[0]
            return (UTFTypeValues.bytes, self.data[lo:lo + raw[1]]), 8
        return (_TYPES[kind], raw[0]), struct.calcsize(fmt)

    def _read(self) -> None:
        at = UTFChunkHeader.size
        per_row, shared = [], {}
        null_cols, const_cols = [], []                                  # (name, the entry the reference's `table` holds)
        for _ in range(self.num_columns):
            flag = self.data[at]
            storage, kind = flag >> 4, flag & 0xF
            if kind > 0xB:
                raise Exception("Unkown data type.")
            name = self._string(struct.unpack_from(">I", self.data, at + 1)[0])
            at += 5
            if storage == 0x1:
                shared[name] = ((UTFTypeValues.string, "<NULL>") if kind == 0xA else
                                (UTFTypeValues.bytes, b"") if kind == 0xB else (_TYPES[kind], None))
                null_cols.append((name, "<NULL>" if kind == 0xA else b"" if kind == 0xB else 0))
            elif storage == 0x3:
                shared[name], used = self._value(kind, at)
                # numeric constants sit in the reference's table as the raw unpack() tuple (utf.py:124)
                const_cols.append((name, shared[name][1] if kind in (0xA, 0xB) else (shared[name][1],)))
                at += used
            elif storage == 0x5:
                per_row.append((name, kind))
            elif storage == 0x7:
                raise NotImplementedError("Unsupported 0x70 storage flag.")
            else:
                raise Exception("Unknown storage flag.")
        self._payload: List[dict] = []
        if not per_row or self.num_rows == 0:
            self._payload.append(dict(shared))
        else:
            for r in range(self.num_rows):
                at = 8 + self.rows_offset + r * self.row_length
                row = {}
                for name, kind in per_row:
                    row[name], used = self._value(kind, at)
                    at += used
                row.update(shared)
                self._payload.append(row)
        # The column-major view, as read_rows_and_columns builds it (utf.py:113-152): columns without storage first
        # (one entry: 0 / "<NULL>" / b""), then constants (one entry), then the per-row columns (one entry per row).
        self.table = {}
        for name, entry in null_cols + const_cols:
            self.table.setdefault(name, []).append(entry)
        if per_row and self.num_rows:
            for row in self._payload:
                for name, _ in per_row:
                    self.table.setdefault(name, []).append(row[name][1])

    def get_payload(self) -> list:
        return self._payload
