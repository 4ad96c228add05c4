"""ACB cue-sheet front-end whose `extract(decode=True)` hands every HCA waveform of the sheet's bank to the batch
decoder in ONE call (the reference decodes them one at a time: PyCriCodecs/acb.py:141-154).

Mirrors the reference class (acb.py:9-47, 141-161): an ACB is one @UTF table whose byte cells may hold further @UTF
tables (parsed recursively into `payload`), with the waveform bank either embedded (`AwbFile`) or next to the sheet as
`<Name>.awb`; `extract` names the files `<index><extension>` with the extension taken from the waveform's
`EncodeType`, and decodes only `.hca` entries.
"""
from __future__ import annotations

import os
from typing import List, Optional

from .awb import AWB
from .chunk import UTFType, UTFTypeValues
from .hca import HCA
from .utf import UTF


class ACB:
    __slots__ = ["filename", "payload", "awb"]

    def __init__(self, filename) -> None:
        self.payload = UTF(filename).get_payload()
        self.filename = filename
        self.acbparse(self.payload)
        self.load_awb()

    def acbparse(self, payload: list) -> None:
        """Replace every byte cell that holds an @UTF table by that table's payload, recursively (acb.py:23-32)."""
        for row in payload:
            for k, v in row.items():
                if isinstance(v, tuple) and v[0] == UTFTypeValues.bytes and v[1].startswith(UTFType.UTF.value):
                    sub = UTF(v[1]).get_payload()
                    self.acbparse(sub)
                    row[k] = sub

    def load_awb(self) -> None:
        """The bank is the `AwbFile` cell, or -- when that is empty -- `<Name>.awb` beside the sheet (acb.py:34-46)."""
        head = self.payload[0]
        embedded = head["AwbFile"][1]
        if embedded == b"":
            name = head["Name"][1] + ".awb"
            self.awb = AWB(os.path.join(os.path.dirname(self.filename), name) if isinstance(self.filename, str) else name)
        else:
            self.awb = AWB(embedded)

    def get_extension(self, EncodeType: int) -> str:
        """acb.py:156-161 (anything that is neither ADX nor HCA falls through to `None` in the reference)."""
        if EncodeType in (0, 3):
            return ".adx"
        if EncodeType in (2, 6):
            return ".hca"
        return None

    def _extensions(self) -> List[Optional[str]]:
        table = self.payload[0]["WaveformTable"]
        return [self.get_extension(table[i]["EncodeType"][1]) for i in range(self.awb.numfiles)]

    def decode_all(self, key: int = 0, ctx=None, raise_errors: bool = True) -> List[Optional[bytes]]:
        """WAV bytes of every waveform the sheet marks as HCA (None for the others), one batch call, bank subkey."""
        files = self.awb.getfiles()
        idx = [i for i, e in enumerate(self._extensions()) if e == ".hca"]
        out: List[Optional[bytes]] = [None] * len(files)
        if idx:
            got = HCA.decode_batch([files[i] for i in idx], keys=key, subkeys=self.awb.subkey, ctx=ctx, raise_errors=raise_errors)
            for i, g in zip(idx, got):
                out[i] = g
        return out

    def extract(self, decode: bool = False, key: int = 0, dirname: str = "", ctx=None) -> List[str]:
        """Write the bank's files as `<index>.wav` (decoded HCA) or `<index><extension>`; returns the paths."""
        if dirname:
            os.makedirs(dirname, exist_ok=True)
        files = self.awb.getfiles()
        exts = self._extensions()
        wavs = self.decode_all(key, ctx) if decode else None
        written = []
        for i, blob in enumerate(files):
            if decode and exts[i] == ".hca":
                path, payload = os.path.join(dirname, f"{i}.wav"), wavs[i]
            else:
                path, payload = os.path.join(dirname, f"{i}{exts[i]}"), blob
            with open(path, "wb") as f:
                f.write(payload)
            written.append(path)
        return written
