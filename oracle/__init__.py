"""TEST INFRASTRUCTURE ONLY: ctypes front-ends of the two checkers.

    oracle.port  -- libcri_oracle.so, the plain-C restatement (cri_oracle.c)
    oracle.ref   -- oracle/_ref/CriCodecs*.so, the UNMODIFIED reference compiled
                    by oracle/Makefile (present when built in the dev container;
                    the prebuilt file travels to the GPU box)

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs import this package. The product (pycricodecs_b200) never does.
"""
from __future__ import annotations

import ctypes
import glob
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_u8p = ctypes.POINTER(ctypes.c_uint8)


def build(ref: bool = True) -> None:
    """Compile the checkers (oracle always; _ref only where /root/reference exists)."""
    subprocess.run(["make", "-s", "-C", HERE, "oracle"], check=True)
    if ref and os.path.isdir("/root/reference/CriCodecs"):
        subprocess.run(["make", "-s", "-C", HERE, "ref"], check=True)


def _buf(b):
    return (ctypes.c_uint8 * len(b)).from_buffer_copy(b)


class _Port:
    def __init__(self):
        path = os.path.join(HERE, "libcri_oracle.so")
        if not os.path.exists(path):
            build(ref=False)
        self.lib = L = ctypes.CDLL(path)
        sz = ctypes.c_size_t
        L.cri_oracle_adx_decode.argtypes = [ctypes.c_char_p, sz, ctypes.c_void_p, sz, ctypes.POINTER(sz)]
        L.cri_oracle_adx_decoded_size.argtypes = [ctypes.c_char_p, sz, ctypes.POINTER(sz)]
        L.cri_oracle_adx_encode.argtypes = [ctypes.c_char_p, sz] + [ctypes.c_uint] * 6 + [ctypes.c_void_p, sz, ctypes.POINTER(sz)]
        L.cri_oracle_hca_decode.argtypes = [ctypes.c_char_p, sz, ctypes.c_uint64, ctypes.c_uint, ctypes.c_void_p, sz, ctypes.POINTER(sz)]
        L.cri_oracle_hca_decoded_size.argtypes = [ctypes.c_char_p, sz, ctypes.POINTER(sz)]
        L.cri_oracle_hca_crypt.argtypes = [ctypes.c_void_p, sz, ctypes.c_int, ctypes.c_uint, ctypes.c_uint64, ctypes.c_uint]
        L.cri_oracle_hca_encode.argtypes = [ctypes.c_char_p, sz, ctypes.c_uint, ctypes.c_void_p, sz, ctypes.POINTER(sz)]
        L.cri_oracle_hca_encoded_size.argtypes = [ctypes.c_char_p, sz, ctypes.c_uint, ctypes.POINTER(sz)]
        L.cri_oracle_crc16.argtypes = [ctypes.c_char_p, sz]
        L.cri_oracle_crc16.restype = ctypes.c_uint
        L.cri_oracle_cipher_table.argtypes = [ctypes.c_int, ctypes.c_uint64, ctypes.c_void_p]
        L.cri_oracle_mix_subkey.argtypes = [ctypes.c_uint64, ctypes.c_uint]
        L.cri_oracle_mix_subkey.restype = ctypes.c_uint64
        L.cri_oracle_hca_decode_range.argtypes = [ctypes.c_char_p, sz, ctypes.c_uint64, ctypes.c_uint, ctypes.c_uint, ctypes.c_void_p]
        L.cri_oracle_hca_unpack.argtypes = [ctypes.c_char_p, sz, ctypes.c_uint64, ctypes.c_uint] + [ctypes.c_void_p] * 5 + [ctypes.POINTER(ctypes.c_int)]
        L.cri_oracle_hca_info.argtypes = [ctypes.c_char_p, sz, ctypes.c_void_p]
        L.cri_oracle_adx_coefficients.argtypes = [ctypes.c_uint, ctypes.c_uint, ctypes.c_void_p]
        for f in ("imdct", "mdct"):
            getattr(L, "cri_oracle_" + f).restype = None

    # -- whole-stream entry points: return (status, bytes) --------------
    def adx_decode(self, adx: bytes):
        n = ctypes.c_size_t(0)
        r = self.lib.cri_oracle_adx_decoded_size(adx, len(adx), ctypes.byref(n))
        if r < 0:
            return r, b""
        out = ctypes.create_string_buffer(n.value)
        r = self.lib.cri_oracle_adx_decode(adx, len(adx), out, n.value, ctypes.byref(n))
        return r, out.raw[:n.value]

    def adx_encode(self, wav: bytes, bitdepth=4, blocksize=18, mode=3, highpass=500, filt=0, version=4):
        n = ctypes.c_size_t(0)
        r = self.lib.cri_oracle_adx_encode(wav, len(wav), bitdepth, blocksize, mode, highpass, filt, version, None, 0, ctypes.byref(n))
        if r < 0:
            return r, b""
        out = ctypes.create_string_buffer(n.value)
        r = self.lib.cri_oracle_adx_encode(wav, len(wav), bitdepth, blocksize, mode, highpass, filt, version, out, n.value, ctypes.byref(n))
        return r, out.raw[:n.value]

    def hca_decode(self, hca: bytes, key=0, subkey=0):
        n = ctypes.c_size_t(0)
        r = self.lib.cri_oracle_hca_decoded_size(hca, len(hca), ctypes.byref(n))
        if r < 0:
            return r, b""
        out = ctypes.create_string_buffer(n.value)
        r = self.lib.cri_oracle_hca_decode(hca, len(hca), key, subkey, out, n.value, ctypes.byref(n))
        return r, out.raw[:n.value]

    def hca_crypt(self, hca: bytes, encrypt: int, ciph_type: int, key: int, subkey=0):
        buf = ctypes.create_string_buffer(bytes(hca), len(hca))
        r = self.lib.cri_oracle_hca_crypt(buf, len(hca), encrypt, ciph_type, key, subkey)
        return r, buf.raw

    def hca_encode(self, wav: bytes, quality=1):
        n = ctypes.c_size_t(0)
        r = self.lib.cri_oracle_hca_encoded_size(wav, len(wav), quality, ctypes.byref(n))
        if r < 0:
            return r, b""
        out = ctypes.create_string_buffer(n.value)
        r = self.lib.cri_oracle_hca_encode(wav, len(wav), quality, out, n.value, ctypes.byref(n))
        return r, out.raw[:n.value]

    # -- probes -----------------------------------------------------------
    def crc16(self, data: bytes) -> int:
        return self.lib.cri_oracle_crc16(data, len(data))

    def cipher_table(self, ciph_type: int, key: int) -> bytes:
        t = ctypes.create_string_buffer(256)
        self.lib.cri_oracle_cipher_table(ciph_type, key, t)
        return t.raw

    def mix_subkey(self, key: int, subkey: int) -> int:
        return self.lib.cri_oracle_mix_subkey(key, subkey)

    def adx_coefficients(self, highpass: int, rate: int):
        c = (ctypes.c_int * 2)()
        self.lib.cri_oracle_adx_coefficients(highpass, rate, c)
        return c[0], c[1]

    def hca_info(self, hca: bytes):
        v = (ctypes.c_uint * 16)()
        r = self.lib.cri_oracle_hca_info(hca, len(hca), v)
        keys = ["version", "header_size", "channels", "rate", "frame_count", "delay", "padding", "frame_size", "total_bands",
                "base_bands", "stereo_bands", "bands_per_hfr", "hfr_groups", "ciph_type", "loop_flag", "min_res"]
        return r, dict(zip(keys, list(v)))

    def hca_decode_range(self, hca: bytes, key: int, f0: int, f1: int, channels: int) -> np.ndarray:
        out = np.zeros(((f1 - f0) * 1024, channels), dtype=np.int16)
        r = self.lib.cri_oracle_hca_decode_range(hca, len(hca), key, f0, f1, out.ctypes.data)
        if r < 0:
            raise ValueError(f"oracle decode_range failed: {r}")
        return out

    def hca_unpack(self, hca: bytes, key: int, frame: int, channels: int):
        sf = np.zeros((channels, 128), np.uint8); res = np.zeros((channels, 128), np.uint8)
        inten = np.zeros((channels, 8), np.uint8); gain = np.zeros((channels, 128), np.float32)
        spec = np.zeros((channels, 8, 128), np.float32); bits = ctypes.c_int(0)
        r = self.lib.cri_oracle_hca_unpack(hca, len(hca), key, frame, sf.ctypes.data, res.ctypes.data, inten.ctypes.data,
                                           gain.ctypes.data, spec.ctypes.data, ctypes.byref(bits))
        return r, dict(sf=sf, res=res, intensity=inten, gain=gain, spectra=spec, bits=bits.value)

    def imdct(self, spectra: np.ndarray, prev: np.ndarray):
        spectra = np.ascontiguousarray(spectra, np.float32); prev = np.array(prev, np.float32)
        wave = np.zeros(128, np.float32); dct = np.zeros(128, np.float32)
        self.lib.cri_oracle_imdct(ctypes.c_void_p(spectra.ctypes.data), ctypes.c_void_p(prev.ctypes.data),
                                  ctypes.c_void_p(wave.ctypes.data), ctypes.c_void_p(dct.ctypes.data))
        return wave, prev, dct

    def mdct(self, wave: np.ndarray, prev: np.ndarray):
        wave = np.ascontiguousarray(wave, np.float32); prev = np.array(prev, np.float32)
        spec = np.zeros(128, np.float32)
        self.lib.cri_oracle_mdct(ctypes.c_void_p(wave.ctypes.data), ctypes.c_void_p(prev.ctypes.data), ctypes.c_void_p(spec.ctypes.data))
        return spec, prev


class _Ref:
    """The compiled reference: its own `CriCodecs` module + the ref_* harness."""

    def __init__(self):
        hits = glob.glob(os.path.join(HERE, "_ref", "CriCodecs*.so"))
        if not hits:
            raise FileNotFoundError("oracle/_ref is not built (make -C oracle ref, needs /root/reference)")
        # load THIS build under its own module object: another `CriCodecs` (the drop-in, or the stock build under
        # baseline/_ref) may already sit in sys.modules, and `import CriCodecs` would silently hand that one back
        import importlib.machinery
        import importlib.util
        loader = importlib.machinery.ExtensionFileLoader("CriCodecs", hits[0])
        spec = importlib.util.spec_from_loader("CriCodecs", loader)
        self.mod = importlib.util.module_from_spec(spec)
        loader.exec_module(self.mod)
        self.lib = ctypes.PyDLL(hits[0])
        self.lib.ref_crc16.restype = ctypes.c_uint
        self.lib.ref_hca_decode_range.argtypes = [ctypes.c_char_p, ctypes.c_uint, ctypes.c_uint64, ctypes.c_uint, ctypes.c_uint, ctypes.c_void_p]
        self.lib.ref_hca_unpack_dump.argtypes = [ctypes.c_char_p, ctypes.c_uint, ctypes.c_uint64, ctypes.c_uint] + [ctypes.c_void_p] * 5 + [ctypes.POINTER(ctypes.c_int)]
        self.lib.ref_cipher_table.argtypes = [ctypes.c_int, ctypes.c_uint64, ctypes.c_void_p]

    def adx_decode(self, adx: bytes) -> bytes:
        return self.mod.AdxDecode(bytes(adx))

    def adx_encode(self, wav: bytes, bitdepth=4, blocksize=18, mode=3, highpass=500, filt=0, version=4, force_noloop=0):
        out = ctypes.create_string_buffer(len(wav) * 2 + 4096)
        n = ctypes.c_uint(0)
        r = self.lib.ref_adx_encode(bytes(wav), bitdepth, blocksize, mode, highpass, filt, version, force_noloop, out, ctypes.byref(n))
        return r, out.raw[:n.value]

    def hca_decode(self, hca: bytes, key=0, subkey=0) -> bytes:
        hs = int.from_bytes(hca[6:8], "big")
        return self.mod.HcaDecode(bytes(hca), hs, key, subkey)

    def hca_crypt(self, hca: bytes, encrypt: int, ciph_type: int, key: int, subkey=0) -> bytes:
        hs = int.from_bytes(hca[6:8], "big")
        return self.mod.HcaCrypt(bytearray(hca), encrypt, hs, ciph_type, key, subkey)  # mutates its argument: hand it a copy

    def hca_encode(self, wav: bytes, quality=1, force_noloop=0):
        n = ctypes.c_uint(0)
        r = self.lib.ref_hca_encode(bytes(wav), quality, force_noloop, None, ctypes.byref(n))
        if r < 0:
            return r, b""
        out = ctypes.create_string_buffer(n.value)
        r = self.lib.ref_hca_encode(bytes(wav), quality, force_noloop, out, ctypes.byref(n))
        return r, out.raw[:n.value]

    def crc16(self, data: bytes) -> int:
        return self.lib.ref_crc16(bytes(data), len(data))

    def cipher_table(self, ciph_type: int, key: int) -> bytes:
        t = ctypes.create_string_buffer(256)
        self.lib.ref_cipher_table(ciph_type, key, t)
        return t.raw

    def adx_coefficients(self, highpass: int, rate: int):
        c = (ctypes.c_int * 2)()
        self.lib.ref_adx_coefficients(highpass, rate, c)
        return c[0], c[1]

    def hca_decode_range(self, hca: bytes, key: int, f0: int, f1: int, channels: int) -> np.ndarray:
        out = np.zeros(((f1 - f0) * 1024, channels), dtype=np.int16)
        r = self.lib.ref_hca_decode_range(bytes(hca), len(hca), key, f0, f1, out.ctypes.data)
        if r < 0:
            raise ValueError(f"reference decode_range failed: {r}")
        return out

    def hca_unpack(self, hca: bytes, key: int, frame: int, channels: int):
        sf = np.zeros((channels, 128), np.uint8); res = np.zeros((channels, 128), np.uint8)
        inten = np.zeros((channels, 8), np.uint8); gain = np.zeros((channels, 128), np.float32)
        spec = np.zeros((channels, 8, 128), np.float32); bits = ctypes.c_int(0)
        r = self.lib.ref_hca_unpack_dump(bytes(hca), len(hca), key, frame, sf.ctypes.data, res.ctypes.data, inten.ctypes.data,
                                         gain.ctypes.data, spec.ctypes.data, ctypes.byref(bits))
        return r, dict(sf=sf, res=res, intensity=inten, gain=gain, spectra=spec, bits=bits.value)

    def imdct(self, spectra: np.ndarray, prev: np.ndarray):
        spectra = np.ascontiguousarray(spectra, np.float32); prev = np.array(prev, np.float32)
        wave = np.zeros(128, np.float32); dct = np.zeros(128, np.float32)
        self.lib.ref_imdct(ctypes.c_void_p(spectra.ctypes.data), ctypes.c_void_p(prev.ctypes.data),
                           ctypes.c_void_p(wave.ctypes.data), ctypes.c_void_p(dct.ctypes.data))
        return wave, prev, dct

    def mdct(self, wave: np.ndarray, prev: np.ndarray):
        wave = np.ascontiguousarray(wave, np.float32); prev = np.array(prev, np.float32)
        spec = np.zeros(128, np.float32)
        self.lib.ref_mdct(ctypes.c_void_p(wave.ctypes.data), ctypes.c_void_p(prev.ctypes.data), ctypes.c_void_p(spec.ctypes.data))
        return spec, prev

    def table(self, name: str, dtype) -> np.ndarray:
        n = self.lib.ref_table(name.encode(), None)
        if n < 0:
            raise KeyError(name)
        buf = ctypes.create_string_buffer(n)
        self.lib.ref_table(name.encode(), buf)
        return np.frombuffer(buf.raw, dtype=dtype).copy()


_port = None
_ref = None


def port() -> _Port:
    global _port
    if _port is None:
        _port = _Port()
    return _port


def ref() -> _Ref:
    global _ref
    if _ref is None:
        _ref = _Ref()
    return _ref


class _Checker:
    """One calling convention over both checkers, output bytes only: the compiled reference where it is built (the GPU box
    gets oracle/_ref with the snapshot), else the C restatement. What the GPU tests and smoke() compare against."""

    def __init__(self):
        self.kind = "reference" if have_ref() else "port"
        self.impl = ref() if have_ref() else port()

    @staticmethod
    def _bytes(r):
        return bytes(r[1]) if isinstance(r, tuple) else bytes(r)

    def hca_decode(self, hca, key=0, subkey=0):
        return self._bytes(self.impl.hca_decode(hca, key, subkey))

    def hca_encode(self, wav, quality=1):
        return self._bytes(self.impl.hca_encode(wav, quality))

    def hca_crypt(self, hca, encrypt, ciph_type, key, subkey=0):
        return self._bytes(self.impl.hca_crypt(hca, encrypt, ciph_type, key, subkey))

    def adx_decode(self, adx):
        return self._bytes(self.impl.adx_decode(adx))

    def adx_encode(self, wav, **kw):
        # always the restatement: ADX::Encode sizes its header from an uninitialised field (adx.cpp:482), so the compiled
        # reference's output depends on what the stack held before the call (seen once in a few hundred runs here); the
        # restatement is pinned against it in tests/test_oracle_golden.py, in a fresh process
        return self._bytes(port().adx_encode(wav, **kw))


_checker = None


def checker() -> _Checker:
    global _checker
    if _checker is None:
        _checker = _Checker()
    return _checker


def have_ref() -> bool:
    return bool(glob.glob(os.path.join(HERE, "_ref", "CriCodecs*.so")))
