"""ctypes windows onto the TEST-ONLY checkers: oracle/liboracle.so (plain-C restatement) and, when present,
oracle/_ref/libsimdutf_ref.so (the unmodified reference compiled from /root/reference).  Imported by tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs only."""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_SO = os.path.join(ORACLE_DIR, "liboracle.so")
REF_SO = os.path.join(ORACLE_DIR, "_ref", "libsimdutf_ref.so")


class Res(ctypes.Structure):
    _fields_ = [("error", ctypes.c_int32), ("count", ctypes.c_uint64)]


class Full(ctypes.Structure):
    _fields_ = [("error", ctypes.c_int32), ("input_count", ctypes.c_uint64), ("output_count", ctypes.c_uint64)]


def _u8(data) -> np.ndarray:
    if isinstance(data, np.ndarray):
        return np.ascontiguousarray(data).view(np.uint8).reshape(-1)
    return np.frombuffer(bytes(data), dtype=np.uint8)


def _u16(data) -> np.ndarray:
    if isinstance(data, np.ndarray):
        return np.ascontiguousarray(data).view(np.uint16).reshape(-1)
    return np.frombuffer(bytes(data), dtype=np.uint16)


def _p(a: np.ndarray):
    return ctypes.c_void_p(a.ctypes.data if a.size else 0)


class Oracle:
    """oracle/oracle.c"""

    def __init__(self):
        if not os.path.exists(ORACLE_SO) or os.path.getmtime(ORACLE_SO) < os.path.getmtime(os.path.join(ORACLE_DIR, "oracle.c")):
            subprocess.check_call(["make", "-s", "-C", ORACLE_DIR, "liboracle.so"])
        L = ctypes.CDLL(ORACLE_SO)
        for f in ("oracle_validate_utf8_with_errors", "oracle_convert_utf8_to_utf16le_with_errors",
                  "oracle_convert_utf8_to_utf32_with_errors", "oracle_convert_utf16le_to_utf8_with_errors",
                  "oracle_validate_utf16le_with_errors", "oracle_base64_to_binary",
                  "oracle_validate_utf16be_with_errors", "oracle_convert_utf16be_to_utf8_with_errors",
                  "oracle_convert_utf8_to_utf16be_with_errors", "oracle_validate_utf32_with_errors",
                  "oracle_convert_utf32_to_utf8_with_errors", "oracle_convert_utf32_to_utf16le_with_errors",
                  "oracle_convert_utf32_to_utf16be_with_errors", "oracle_convert_utf16le_to_utf32_with_errors",
                  "oracle_convert_utf16be_to_utf32_with_errors", "oracle_validate_ascii_with_errors",
                  "oracle_convert_utf8_to_latin1_with_errors", "oracle_convert_utf16_to_latin1_with_errors",
                  "oracle_convert_utf32_to_latin1_with_errors"):
            getattr(L, f).restype = Res
        L.oracle_base64_to_binary_details.restype = Full
        for f in ("oracle_count_utf8", "oracle_utf16_length_from_utf8", "oracle_utf32_length_from_utf8",
                  "oracle_count_utf16le", "oracle_utf8_length_from_utf16le", "oracle_utf32_length_from_utf16le",
                  "oracle_maximal_binary_length_from_base64", "oracle_trim_partial_utf8", "oracle_trim_partial_utf16le",
                  "oracle_base64_length_from_binary", "oracle_binary_to_base64", "oracle_count_utf16be",
                  "oracle_utf8_length_from_utf16be", "oracle_utf32_length_from_utf16be", "oracle_utf8_length_from_utf32",
                  "oracle_utf16_length_from_utf32", "oracle_utf8_length_from_latin1", "oracle_convert_latin1_to_utf8",
                  "oracle_convert_latin1_to_utf16", "oracle_convert_latin1_to_utf32"):
            getattr(L, f).restype = ctypes.c_uint64
        self.L = L

    # --- UTF-8 ---
    def validate_utf8_with_errors(self, data):
        a = _u8(data)
        r = self.L.oracle_validate_utf8_with_errors(_p(a), ctypes.c_size_t(a.size))
        return (r.error, r.count)

    def count_utf8(self, data):
        a = _u8(data)
        return int(self.L.oracle_count_utf8(_p(a), ctypes.c_size_t(a.size)))

    def utf16_length_from_utf8(self, data):
        a = _u8(data)
        return int(self.L.oracle_utf16_length_from_utf8(_p(a), ctypes.c_size_t(a.size)))

    def convert_utf8_to_utf16le_with_errors(self, data):
        a = _u8(data)
        out = np.zeros(a.size + 8, dtype=np.uint16)
        r = self.L.oracle_convert_utf8_to_utf16le_with_errors(_p(a), ctypes.c_size_t(a.size), _p(out))
        return (r.error, r.count), (out[: r.count] if r.error == 0 else out[:0])

    def convert_utf8_to_utf32_with_errors(self, data):
        a = _u8(data)
        out = np.zeros(a.size + 8, dtype=np.uint32)
        r = self.L.oracle_convert_utf8_to_utf32_with_errors(_p(a), ctypes.c_size_t(a.size), _p(out))
        return (r.error, r.count), (out[: r.count] if r.error == 0 else out[:0])

    # --- UTF-16 ---
    def validate_utf16le_with_errors(self, data):
        a = _u16(data)
        r = self.L.oracle_validate_utf16le_with_errors(_p(a), ctypes.c_size_t(a.size))
        return (r.error, r.count)

    def count_utf16le(self, data):
        a = _u16(data)
        return int(self.L.oracle_count_utf16le(_p(a), ctypes.c_size_t(a.size)))

    def utf8_length_from_utf16le(self, data):
        a = _u16(data)
        return int(self.L.oracle_utf8_length_from_utf16le(_p(a), ctypes.c_size_t(a.size)))

    def convert_utf16le_to_utf8_with_errors(self, data):
        a = _u16(data)
        out = np.zeros(3 * a.size + 8, dtype=np.uint8)
        r = self.L.oracle_convert_utf16le_to_utf8_with_errors(_p(a), ctypes.c_size_t(a.size), _p(out))
        return (r.error, r.count), (out[: r.count] if r.error == 0 else out[:0])

    # --- UTF-16BE twins: arrays hold big-endian units (i.e. byte-swapped values on this little-endian host) ---
    def validate_utf16be_with_errors(self, data):
        a = _u16(data)
        r = self.L.oracle_validate_utf16be_with_errors(_p(a), ctypes.c_size_t(a.size))
        return (r.error, r.count)

    def count_utf16be(self, data):
        a = _u16(data)
        return int(self.L.oracle_count_utf16be(_p(a), ctypes.c_size_t(a.size)))

    def utf8_length_from_utf16be(self, data):
        a = _u16(data)
        return int(self.L.oracle_utf8_length_from_utf16be(_p(a), ctypes.c_size_t(a.size)))

    def convert_utf16be_to_utf8_with_errors(self, data):
        a = _u16(data)
        out = np.zeros(3 * a.size + 8, dtype=np.uint8)
        r = self.L.oracle_convert_utf16be_to_utf8_with_errors(_p(a), ctypes.c_size_t(a.size), _p(out))
        return (r.error, r.count), (out[: r.count] if r.error == 0 else out[:0])

    def convert_utf8_to_utf16be_with_errors(self, data):
        a = _u8(data)
        out = np.zeros(a.size + 8, dtype=np.uint16)
        r = self.L.oracle_convert_utf8_to_utf16be_with_errors(_p(a), ctypes.c_size_t(a.size), _p(out))
        return (r.error, r.count), (out[: r.count] if r.error == 0 else out[:0])

    def change_endianness_utf16(self, data):
        a = _u16(data)
        out = np.zeros(a.size, dtype=np.uint16)
        self.L.oracle_change_endianness_utf16(_p(a), ctypes.c_size_t(a.size), _p(out))
        return out

    # --- UTF-32 family ---
    def validate_utf32_with_errors(self, data):
        a = np.ascontiguousarray(data, dtype=np.uint32)
        r = self.L.oracle_validate_utf32_with_errors(_p(a), ctypes.c_size_t(a.size))
        return (r.error, r.count)

    def utf8_length_from_utf32(self, data):
        a = np.ascontiguousarray(data, dtype=np.uint32)
        return int(self.L.oracle_utf8_length_from_utf32(_p(a), ctypes.c_size_t(a.size)))

    def utf16_length_from_utf32(self, data):
        a = np.ascontiguousarray(data, dtype=np.uint32)
        return int(self.L.oracle_utf16_length_from_utf32(_p(a), ctypes.c_size_t(a.size)))

    def _conv(self, fn, a, out):
        r = getattr(self.L, fn)(_p(a), ctypes.c_size_t(a.size), _p(out))
        return (r.error, r.count), (out[: r.count] if r.error == 0 else out[:0])

    def convert_utf32_to_utf8_with_errors(self, data):
        a = np.ascontiguousarray(data, dtype=np.uint32)
        return self._conv("oracle_convert_utf32_to_utf8_with_errors", a, np.zeros(4 * a.size + 8, dtype=np.uint8))

    def convert_utf32_to_utf16_with_errors(self, data, be=False):
        a = np.ascontiguousarray(data, dtype=np.uint32)
        return self._conv("oracle_convert_utf32_to_utf16be_with_errors" if be else "oracle_convert_utf32_to_utf16le_with_errors",
                          a, np.zeros(2 * a.size + 8, dtype=np.uint16))

    def convert_utf16_to_utf32_with_errors(self, data, be=False):
        a = _u16(data)
        return self._conv("oracle_convert_utf16be_to_utf32_with_errors" if be else "oracle_convert_utf16le_to_utf32_with_errors",
                          a, np.zeros(a.size + 8, dtype=np.uint32))

    # --- Latin-1 / ASCII family ---
    def validate_ascii_with_errors(self, data):
        a = _u8(data)
        r = self.L.oracle_validate_ascii_with_errors(_p(a), ctypes.c_size_t(a.size))
        return (r.error, r.count)

    def utf8_length_from_latin1(self, data):
        a = _u8(data)
        return int(self.L.oracle_utf8_length_from_latin1(_p(a), ctypes.c_size_t(a.size)))

    def convert_latin1_to_utf8(self, data):
        a = _u8(data); out = np.zeros(2 * a.size + 8, dtype=np.uint8)
        n = int(self.L.oracle_convert_latin1_to_utf8(_p(a), ctypes.c_size_t(a.size), _p(out)))
        return out[:n]

    def convert_latin1_to_utf16(self, data, be=False):
        a = _u8(data); out = np.zeros(a.size + 8, dtype=np.uint16)
        n = int(self.L.oracle_convert_latin1_to_utf16(_p(a), ctypes.c_size_t(a.size), _p(out), int(be)))
        return out[:n]

    def convert_latin1_to_utf32(self, data):
        a = _u8(data); out = np.zeros(a.size + 8, dtype=np.uint32)
        n = int(self.L.oracle_convert_latin1_to_utf32(_p(a), ctypes.c_size_t(a.size), _p(out)))
        return out[:n]

    def convert_utf8_to_latin1_with_errors(self, data):
        a = _u8(data)
        return self._conv("oracle_convert_utf8_to_latin1_with_errors", a, np.zeros(a.size + 8, dtype=np.uint8))

    def convert_utf16_to_latin1_with_errors(self, data, be=False):
        a = _u16(data); out = np.zeros(a.size + 8, dtype=np.uint8)
        r = self.L.oracle_convert_utf16_to_latin1_with_errors(_p(a), ctypes.c_size_t(a.size), _p(out), int(be))
        return (r.error, r.count), (out[: r.count] if r.error == 0 else out[:0])

    def convert_utf32_to_latin1_with_errors(self, data):
        a = np.ascontiguousarray(data, dtype=np.uint32)
        return self._conv("oracle_convert_utf32_to_latin1_with_errors", a, np.zeros(a.size + 8, dtype=np.uint8))

    # --- SURVEY.md §8f rank 4 ---
    def to_well_formed_utf16(self, data, be=False):
        a = _u16(data); out = np.zeros(a.size, dtype=np.uint16)
        self.L.oracle_to_well_formed_utf16(_p(a), ctypes.c_size_t(a.size), _p(out), int(be))
        return out

    def detect_encodings(self, data):
        a = np.zeros(len(bytes(data)) + 4, dtype=np.uint32).view(np.uint8)[: len(bytes(data))]  # 4-byte aligned copy
        a[:] = np.frombuffer(bytes(data), dtype=np.uint8)
        return int(self.L.oracle_detect_encodings(_p(a), ctypes.c_size_t(a.size)))

    # --- base64 ---
    def maximal_binary_length_from_base64(self, data):
        a = _u8(data)
        return int(self.L.oracle_maximal_binary_length_from_base64(_p(a), ctypes.c_size_t(a.size)))

    def base64_to_binary_details(self, data, options=0, last_chunk=0):
        a = _u8(data)
        out = np.zeros(a.size + 8, dtype=np.uint8)
        r = self.L.oracle_base64_to_binary_details(_p(a), ctypes.c_size_t(a.size), _p(out), ctypes.c_uint64(options),
                                                   ctypes.c_uint64(last_chunk))
        return (r.error, r.input_count, r.output_count), out[: r.output_count]

    def binary_to_base64(self, data, options=0):
        a = _u8(data)
        n = int(self.L.oracle_base64_length_from_binary(ctypes.c_size_t(a.size), ctypes.c_uint64(options)))
        out = np.zeros(n + 8, dtype=np.uint8)
        w = int(self.L.oracle_binary_to_base64(_p(a), ctypes.c_size_t(a.size), _p(out), ctypes.c_uint64(options)))
        return out[:w].tobytes()

    def trim_partial_utf8(self, data):
        a = _u8(data)
        return int(self.L.oracle_trim_partial_utf8(_p(a), ctypes.c_size_t(a.size)))


class Reference:
    """oracle/_ref/libsimdutf_ref.so — the unmodified reference (icelake / haswell / westmere / fallback)."""

    @staticmethod
    def load_or_none():
        if not os.path.exists(REF_SO):
            return None
        try:
            return Reference()
        except OSError:
            return None

    def __init__(self):
        L = ctypes.CDLL(REF_SO)
        L.ref_best_name.restype = ctypes.c_char_p
        for f in ("ref_count_utf8", "ref_utf16_length_from_utf8", "ref_utf32_length_from_utf8", "ref_count_utf16le",
                  "ref_utf8_length_from_utf16le", "ref_maximal_binary_length_from_base64", "ref_convert_utf8_to_utf16le",
                  "ref_convert_utf8_to_utf32", "ref_convert_utf16le_to_utf8", "ref_binary_to_base64",
                  "ref_base64_length_from_binary", "ref_trim_partial_utf8",
                  "ref_mt_utf16_length_then_convert_utf8_to_utf16le"):
            getattr(L, f).restype = ctypes.c_int64
        self.L = L
        self.best = L.ref_best_name().decode()

    def impls(self):
        return [n for n in ("icelake", "haswell", "westmere", "fallback") if self.L.ref_has_impl(n.encode())]

    def validate_utf8_with_errors(self, impl, data):
        a = _u8(data); r = Res()
        assert self.L.ref_validate_utf8_with_errors(impl.encode(), _p(a), ctypes.c_size_t(a.size), ctypes.byref(r)) == 0
        return (r.error, r.count)

    def count_utf8(self, impl, data):
        a = _u8(data)
        return int(self.L.ref_count_utf8(impl.encode(), _p(a), ctypes.c_size_t(a.size)))

    def utf16_length_from_utf8(self, impl, data):
        a = _u8(data)
        return int(self.L.ref_utf16_length_from_utf8(impl.encode(), _p(a), ctypes.c_size_t(a.size)))

    def convert_utf8_to_utf16le_with_errors(self, impl, data):
        a = _u8(data); r = Res()
        out = np.zeros(a.size + 64, dtype=np.uint16)
        assert self.L.ref_convert_utf8_to_utf16le_with_errors(impl.encode(), _p(a), ctypes.c_size_t(a.size), _p(out), ctypes.byref(r)) == 0
        return (r.error, r.count), (out[: r.count] if r.error == 0 else out[:0])

    def convert_utf8_to_utf32_with_errors(self, impl, data):
        a = _u8(data); r = Res()
        out = np.zeros(a.size + 64, dtype=np.uint32)
        assert self.L.ref_convert_utf8_to_utf32_with_errors(impl.encode(), _p(a), ctypes.c_size_t(a.size), _p(out), ctypes.byref(r)) == 0
        return (r.error, r.count), (out[: r.count] if r.error == 0 else out[:0])

    def count_utf16le(self, impl, data):
        a = _u16(data)
        return int(self.L.ref_count_utf16le(impl.encode(), _p(a), ctypes.c_size_t(a.size)))

    def utf8_length_from_utf16le(self, impl, data):
        a = _u16(data)
        return int(self.L.ref_utf8_length_from_utf16le(impl.encode(), _p(a), ctypes.c_size_t(a.size)))

    def validate_utf16le_with_errors(self, impl, data):
        a = _u16(data); r = Res()
        assert self.L.ref_validate_utf16le_with_errors(impl.encode(), _p(a), ctypes.c_size_t(a.size), ctypes.byref(r)) == 0
        return (r.error, r.count)

    def convert_utf16le_to_utf8_with_errors(self, impl, data):
        a = _u16(data); r = Res()
        out = np.zeros(3 * a.size + 64, dtype=np.uint8)
        assert self.L.ref_convert_utf16le_to_utf8_with_errors(impl.encode(), _p(a), ctypes.c_size_t(a.size), _p(out), ctypes.byref(r)) == 0
        return (r.error, r.count), (out[: r.count] if r.error == 0 else out[:0])

    def has_be(self):
        return hasattr(self.L, "ref_count_utf16be")

    def has_utf32(self):
        return hasattr(self.L, "ref_validate_utf32_with_errors")

    def has_latin1(self):
        return hasattr(self.L, "ref_validate_ascii_with_errors")

    def has_rank4(self):
        return hasattr(self.L, "ref_detect_encodings")

    def to_well_formed_utf16(self, impl, data, be=False):
        a = _u16(data); out = np.zeros(a.size, dtype=np.uint16)
        assert self.L.ref_to_well_formed_utf16(impl.encode(), int(be), _p(a), ctypes.c_size_t(a.size), _p(out)) == 0
        return out

    def detect_encodings(self, impl, data):
        a = np.zeros(len(bytes(data)) + 4, dtype=np.uint32).view(np.uint8)[: len(bytes(data))]
        a[:] = np.frombuffer(bytes(data), dtype=np.uint8)
        return int(self.L.ref_detect_encodings(impl.encode(), _p(a), ctypes.c_size_t(a.size)))

    def validate_ascii_with_errors(self, impl, data):
        a = _u8(data); r = Res()
        assert self.L.ref_validate_ascii_with_errors(impl.encode(), _p(a), ctypes.c_size_t(a.size), ctypes.byref(r)) == 0
        return (r.error, r.count)

    def utf8_length_from_latin1(self, impl, data):
        a = _u8(data)
        return int(self.L.ref_utf8_length_from_latin1(impl.encode(), _p(a), ctypes.c_size_t(a.size)))

    def latin1_length_from_utf8(self, impl, data):
        a = _u8(data)
        return int(self.L.ref_latin1_length_from_utf8(impl.encode(), _p(a), ctypes.c_size_t(a.size)))

    def convert_latin1_to_utf8(self, impl, data):
        a = _u8(data); out = np.zeros(2 * a.size + 64, dtype=np.uint8)
        n = int(self.L.ref_convert_latin1_to_utf8(impl.encode(), _p(a), ctypes.c_size_t(a.size), _p(out)))
        return out[:n]

    def convert_latin1_to_utf16(self, impl, data, be=False):
        a = _u8(data); out = np.zeros(a.size + 64, dtype=np.uint16)
        n = int(self.L.ref_convert_latin1_to_utf16(impl.encode(), int(be), _p(a), ctypes.c_size_t(a.size), _p(out)))
        return out[:n]

    def convert_latin1_to_utf32(self, impl, data):
        a = _u8(data); out = np.zeros(a.size + 64, dtype=np.uint32)
        n = int(self.L.ref_convert_latin1_to_utf32(impl.encode(), _p(a), ctypes.c_size_t(a.size), _p(out)))
        return out[:n]

    def convert_utf8_to_latin1_with_errors(self, impl, data):
        a = _u8(data); r = Res(); out = np.zeros(a.size + 64, dtype=np.uint8)
        assert self.L.ref_convert_utf8_to_latin1_with_errors(impl.encode(), _p(a), ctypes.c_size_t(a.size), _p(out), ctypes.byref(r)) == 0
        return (r.error, r.count), (out[: r.count] if r.error == 0 else out[:0])

    def convert_utf16_to_latin1_with_errors(self, impl, data, be=False):
        a = _u16(data); r = Res(); out = np.zeros(a.size + 64, dtype=np.uint8)
        assert self.L.ref_convert_utf16_to_latin1_with_errors(impl.encode(), int(be), _p(a), ctypes.c_size_t(a.size), _p(out), ctypes.byref(r)) == 0
        return (r.error, r.count), (out[: r.count] if r.error == 0 else out[:0])

    def convert_utf32_to_latin1_with_errors(self, impl, data):
        a = np.ascontiguousarray(data, dtype=np.uint32); r = Res(); out = np.zeros(a.size + 64, dtype=np.uint8)
        assert self.L.ref_convert_utf32_to_latin1_with_errors(impl.encode(), _p(a), ctypes.c_size_t(a.size), _p(out), ctypes.byref(r)) == 0
        return (r.error, r.count), (out[: r.count] if r.error == 0 else out[:0])

    def validate_utf32_with_errors(self, impl, data):
        a = np.ascontiguousarray(data, dtype=np.uint32); r = Res()
        assert self.L.ref_validate_utf32_with_errors(impl.encode(), _p(a), ctypes.c_size_t(a.size), ctypes.byref(r)) == 0
        return (r.error, r.count)

    def utf8_length_from_utf32(self, impl, data):
        a = np.ascontiguousarray(data, dtype=np.uint32)
        return int(self.L.ref_utf8_length_from_utf32(impl.encode(), _p(a), ctypes.c_size_t(a.size)))

    def utf16_length_from_utf32(self, impl, data):
        a = np.ascontiguousarray(data, dtype=np.uint32)
        return int(self.L.ref_utf16_length_from_utf32(impl.encode(), _p(a), ctypes.c_size_t(a.size)))

    def convert_utf32_to_utf8_with_errors(self, impl, data):
        a = np.ascontiguousarray(data, dtype=np.uint32); r = Res()
        out = np.zeros(4 * a.size + 64, dtype=np.uint8)
        assert self.L.ref_convert_utf32_to_utf8_with_errors(impl.encode(), _p(a), ctypes.c_size_t(a.size), _p(out), ctypes.byref(r)) == 0
        return (r.error, r.count), (out[: r.count] if r.error == 0 else out[:0])

    def convert_utf32_to_utf16_with_errors(self, impl, data, be=False):
        a = np.ascontiguousarray(data, dtype=np.uint32); r = Res()
        out = np.zeros(2 * a.size + 64, dtype=np.uint16)
        assert self.L.ref_convert_utf32_to_utf16_with_errors(impl.encode(), int(be), _p(a), ctypes.c_size_t(a.size), _p(out), ctypes.byref(r)) == 0
        return (r.error, r.count), (out[: r.count] if r.error == 0 else out[:0])

    def convert_utf16_to_utf32_with_errors(self, impl, data, be=False):
        a = _u16(data); r = Res()
        out = np.zeros(a.size + 64, dtype=np.uint32)
        assert self.L.ref_convert_utf16_to_utf32_with_errors(impl.encode(), int(be), _p(a), ctypes.c_size_t(a.size), _p(out), ctypes.byref(r)) == 0
        return (r.error, r.count), (out[: r.count] if r.error == 0 else out[:0])

    def count_utf16be(self, impl, data):
        a = _u16(data)
        return int(self.L.ref_count_utf16be(impl.encode(), _p(a), ctypes.c_size_t(a.size)))

    def utf8_length_from_utf16be(self, impl, data):
        a = _u16(data)
        return int(self.L.ref_utf8_length_from_utf16be(impl.encode(), _p(a), ctypes.c_size_t(a.size)))

    def validate_utf16be_with_errors(self, impl, data):
        a = _u16(data); r = Res()
        assert self.L.ref_validate_utf16be_with_errors(impl.encode(), _p(a), ctypes.c_size_t(a.size), ctypes.byref(r)) == 0
        return (r.error, r.count)

    def convert_utf16be_to_utf8_with_errors(self, impl, data):
        a = _u16(data); r = Res()
        out = np.zeros(3 * a.size + 64, dtype=np.uint8)
        assert self.L.ref_convert_utf16be_to_utf8_with_errors(impl.encode(), _p(a), ctypes.c_size_t(a.size), _p(out), ctypes.byref(r)) == 0
        return (r.error, r.count), (out[: r.count] if r.error == 0 else out[:0])

    def convert_utf8_to_utf16be_with_errors(self, impl, data):
        a = _u8(data); r = Res()
        out = np.zeros(a.size + 64, dtype=np.uint16)
        assert self.L.ref_convert_utf8_to_utf16be_with_errors(impl.encode(), _p(a), ctypes.c_size_t(a.size), _p(out), ctypes.byref(r)) == 0
        return (r.error, r.count), (out[: r.count] if r.error == 0 else out[:0])

    def change_endianness_utf16(self, impl, data):
        a = _u16(data)
        out = np.zeros(a.size, dtype=np.uint16)
        assert self.L.ref_change_endianness_utf16(impl.encode(), _p(a), ctypes.c_size_t(a.size), _p(out)) == 0
        return out

    def base64_to_binary_details(self, impl, data, options=0, last_chunk=0):
        a = _u8(data); r = Full()
        out = np.zeros(a.size + 64, dtype=np.uint8)
        assert self.L.ref_base64_to_binary_details(impl.encode(), _p(a), ctypes.c_size_t(a.size), _p(out), ctypes.c_uint64(options),
                                                   ctypes.c_uint64(last_chunk), ctypes.byref(r)) == 0
        return (r.error, r.input_count, r.output_count), out[: r.output_count]

    def maximal_binary_length_from_base64(self, data):
        a = _u8(data)
        return int(self.L.ref_maximal_binary_length_from_base64(_p(a), ctypes.c_size_t(a.size)))
