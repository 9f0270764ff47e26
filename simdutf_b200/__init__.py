"""simdutf_b200 — Python plumbing over the C ABI of the B200 backend (include/simdutf_b200.h).

The product is the shared library `libsimdutf_b200.so` (sm_100a CUDA kernels + C ABI) and the C++
`simdutf::b200::implementation` subclass built on it (csrc/b200_implementation.cpp).  This module only
loads that library with ctypes so that tests and bench.py can drive the very same entry points the C++
virtuals call.  There is no Python or CPU implementation behind these functions: if the library is
missing, or no sm_100 device is usable, calls fail loudly.

Naming mirrors the reference's free functions (reference include/simdutf/implementation.h:89-3293):
    validate_utf8_with_errors, count_utf8, utf16_length_from_utf8, convert_utf8_to_utf16le_with_errors,
    convert_utf8_to_utf32_with_errors, count_utf16le, utf8_length_from_utf16le,
    convert_utf16le_to_utf8_with_errors, base64_to_binary_details, ...
Each accepts either a CUDA torch tensor (device path, zero copy) or bytes / numpy (host path).
"""
from __future__ import annotations

import ctypes
import os
from dataclasses import dataclass

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG, "libsimdutf_b200.so")

# simdutf::error_code (reference include/simdutf/error.h:5-32)
SUCCESS, HEADER_BITS, TOO_SHORT, TOO_LONG, OVERLONG, TOO_LARGE, SURROGATE = range(7)
INVALID_BASE64_CHARACTER, BASE64_INPUT_REMAINDER, BASE64_EXTRA_BITS, OUTPUT_BUFFER_TOO_SMALL, OTHER = range(7, 12)
ERROR_NAMES = [
    "SUCCESS", "HEADER_BITS", "TOO_SHORT", "TOO_LONG", "OVERLONG", "TOO_LARGE", "SURROGATE",
    "INVALID_BASE64_CHARACTER", "BASE64_INPUT_REMAINDER", "BASE64_EXTRA_BITS", "OUTPUT_BUFFER_TOO_SMALL", "OTHER",
]
# base64_options / last_chunk_handling_options (reference include/simdutf/implementation.h:2782-2811)
base64_default, base64_url, base64_reverse_padding = 0, 1, 2
base64_default_no_padding, base64_url_with_padding = 2, 3
base64_default_accept_garbage, base64_url_accept_garbage = 4, 5
base64_default_or_url, base64_default_or_url_accept_garbage = 8, 12
loose, strict, stop_before_partial = 0, 1, 2


class Result(ctypes.Structure):
    """simdutf::result (reference include/simdutf/error.h:34-37)."""
    _fields_ = [("error", ctypes.c_int32), ("reserved_", ctypes.c_uint32), ("count", ctypes.c_uint64)]

    def astuple(self):
        return (int(self.error), int(self.count))


class FullResult(ctypes.Structure):
    """simdutf::full_result (reference include/simdutf/error.h:54-57)."""
    _fields_ = [("error", ctypes.c_int32), ("reserved_", ctypes.c_uint32),
                ("input_count", ctypes.c_uint64), ("output_count", ctypes.c_uint64)]

    def astuple(self):
        return (int(self.error), int(self.input_count), int(self.output_count))

    def as_result(self):
        """full_result -> result conversion (reference include/simdutf/error.h:66-73)."""
        if self.error in (SUCCESS, BASE64_INPUT_REMAINDER):
            return (int(self.error), int(self.output_count))
        return (int(self.error), int(self.input_count))


class Shard(ctypes.Structure):
    """b200_shard (include/simdutf_b200.h): one device-resident shard of a sharded multi-device call."""
    _fields_ = [("device", ctypes.c_int32), ("reserved_", ctypes.c_uint32), ("d_in", ctypes.c_void_p),
                ("len", ctypes.c_uint64), ("d_out", ctypes.c_void_p)]


class ShardedResultC(ctypes.Structure):
    """b200_sharded_result (include/simdutf_b200.h)."""
    _fields_ = [("error", ctypes.c_int32), ("reserved_", ctypes.c_uint32), ("count", ctypes.c_uint64),
                ("in_offset", ctypes.c_uint64), ("out_offset", ctypes.c_uint64)]

    def astuple(self):
        return (int(self.error), int(self.count), int(self.in_offset), int(self.out_offset))


class B200Error(RuntimeError):
    pass


# Every symbol include/simdutf_b200.h declares: name -> (restype, argtypes)
_vp, _sz, _u64 = ctypes.c_void_p, ctypes.c_size_t, ctypes.c_uint64
_pres, _pfull, _pu64 = ctypes.POINTER(Result), ctypes.POINTER(FullResult), ctypes.POINTER(ctypes.c_uint64)
_I = ctypes.c_int
SYMBOLS = {
    "b200_device_count": (_I, []),
    "b200_set_device": (_I, [_I]),
    "b200_get_device": (_I, []),
    "b200_name": (ctypes.c_char_p, []),
    "b200_description": (ctypes.c_char_p, []),
    "b200_launch_count": (_u64, []),
    "b200_last_error": (ctypes.c_char_p, []),
    "b200_host_alloc": (_I, [ctypes.POINTER(_vp), _sz]),
    "b200_host_free": (_I, [_vp]),
    "b200_host_maximal_binary_length_from_base64": (_sz, [_vp, _sz]),
    "b200_host_trim_partial_utf8": (_sz, [_vp, _sz]),
    "b200_host_trim_partial_utf16le": (_sz, [_vp, _sz]),
    "b200_trim_partial_utf8": (_I, [_vp, _sz, ctypes.POINTER(_sz), _vp]),
}
for _name, _res in [("validate_utf8_with_errors", _pres), ("count_utf8", _pu64), ("utf16_length_from_utf8", _pu64),
                    ("count_utf16le", _pu64), ("utf8_length_from_utf16le", _pu64),
                    ("validate_utf16le_with_errors", _pres), ("count_utf16be", _pu64), ("utf8_length_from_utf16be", _pu64),
                    ("validate_utf16be_with_errors", _pres), ("validate_utf32_with_errors", _pres),
                    ("utf8_length_from_utf32", _pu64), ("utf16_length_from_utf32", _pu64),
                    ("validate_ascii_with_errors", _pres), ("utf8_length_from_latin1", _pu64),
                    ("detect_encodings", _pu64)]:
    SYMBOLS[f"b200_{_name}_async"] = (_I, [_vp, _sz, _vp, _vp])
    SYMBOLS[f"b200_{_name}"] = (_I, [_vp, _sz, _res, _vp])
    SYMBOLS[f"b200_host_{_name}"] = (_I, [_vp, _sz, _res])
for _name in ["convert_utf8_to_utf16le", "convert_utf8_to_utf32", "convert_utf16le_to_utf8", "convert_utf8_to_utf16be",
              "convert_utf16be_to_utf8", "change_endianness_utf16", "convert_utf32_to_utf8", "convert_utf32_to_utf16le",
              "convert_utf32_to_utf16be", "convert_utf16le_to_utf32", "convert_utf16be_to_utf32",
              "convert_latin1_to_utf8", "convert_latin1_to_utf16le", "convert_latin1_to_utf16be", "convert_latin1_to_utf32",
              "convert_utf8_to_latin1", "convert_utf16le_to_latin1", "convert_utf16be_to_latin1", "convert_utf32_to_latin1",
              "to_well_formed_utf16le", "to_well_formed_utf16be"]:
    SYMBOLS[f"b200_{_name}_async"] = (_I, [_vp, _sz, _vp, _vp, _vp])
    SYMBOLS[f"b200_{_name}"] = (_I, [_vp, _sz, _vp, _pres, _vp])
    SYMBOLS[f"b200_host_{_name}"] = (_I, [_vp, _sz, _vp, _pres])
SYMBOLS["b200_base64_to_binary_utf16_async"] = (_I, [_vp, _sz, _vp, _u64, _u64, _vp, _vp])
SYMBOLS["b200_base64_to_binary_utf16"] = (_I, [_vp, _sz, _vp, _u64, _u64, _pfull, _vp])
SYMBOLS["b200_host_base64_to_binary_utf16"] = (_I, [_vp, _sz, _vp, _u64, _u64, _pfull])
SYMBOLS["b200_binary_to_base64_async"] = (_I, [_vp, _sz, _vp, _u64, _vp, _vp])
SYMBOLS["b200_binary_to_base64"] = (_I, [_vp, _sz, _vp, _u64, _pres, _vp])
SYMBOLS["b200_host_binary_to_base64"] = (_I, [_vp, _sz, _vp, _u64, _pres])
SYMBOLS["b200_base64_length_from_binary"] = (_sz, [_sz, _u64])
SYMBOLS["b200_base64_to_binary_async"] = (_I, [_vp, _sz, _vp, _u64, _u64, _vp, _vp])
SYMBOLS["b200_base64_to_binary"] = (_I, [_vp, _sz, _vp, _u64, _u64, _pfull, _vp])
SYMBOLS["b200_host_base64_to_binary"] = (_I, [_vp, _sz, _vp, _u64, _u64, _pfull])
SYMBOLS["b200_sharded_combine_async"] = (_I, [_vp, _I, _I, _I, _vp, _vp])
for _name in ["validate_utf8_with_errors", "utf16_length_from_utf8", "convert_utf8_to_utf16le", "convert_utf8_to_utf32",
              "convert_utf16le_to_utf8"]:
    SYMBOLS[f"b200_mgpu_{_name}"] = (_I, [_vp, _I, _vp])
SYMBOLS["b200_mgpu_last_gather"] = (_I, [])
for _name in ["validate_utf8", "count_utf8", "utf16_length_from_utf8"]:
    SYMBOLS[f"b200_{_name}_batch_async"] = (_I, [_vp, _vp, _sz, _vp, _vp])
    SYMBOLS[f"b200_host_{_name}_batch"] = (_I, [_vp, _vp, _sz, _vp])
for _name in ["convert_utf8_to_utf16le", "convert_utf8_to_utf16be"]:
    SYMBOLS[f"b200_{_name}_batch_async"] = (_I, [_vp, _vp, _sz, _vp, _vp, _vp, _vp])
SYMBOLS["b200_host_convert_utf8_to_utf16le_batch"] = (_I, [_vp, _vp, _sz, _vp, _vp])
SYMBOLS["b200_host_set_devices"] = (_I, [_I])
SYMBOLS["b200_host_get_devices"] = (_I, [])
SYMBOLS["b200_set_tuning"] = (_I, [ctypes.c_char_p, _I])

_lib = None


def load() -> ctypes.CDLL:
    """Load libsimdutf_b200.so (built in-tree by simdutf_b200.build / __graft_entry__.build())."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise B200Error(f"{LIB_PATH} is missing: run `python -m simdutf_b200.build` (there is no fallback path)")
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(lib, name)  # AttributeError if the library does not export a declared symbol
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def _check(status: int, what: str) -> None:
    if status != 0:
        msg = load().b200_last_error().decode("utf-8", "replace")
        raise B200Error(f"{what} failed with status {status}: {msg}")


def device_count() -> int:
    return int(load().b200_device_count())


def set_device(index: int) -> None:
    _check(load().b200_set_device(index), "b200_set_device")


def host_set_devices(n: int) -> None:
    """The calling thread's b200_host_* calls spread large buffers over `n` devices."""
    _check(load().b200_host_set_devices(n), "b200_host_set_devices")


def set_tuning(name: str, value: int) -> None:
    _check(load().b200_set_tuning(name.encode(), value), "b200_set_tuning")


def mgpu(name: str, shards):
    """Sharded multi-device call b200_mgpu_<name>: `shards` = [(device index, input CUDA tensor, output CUDA tensor or
    None), ...] in buffer order.  Returns one (error, count, in_offset, out_offset) per shard."""
    lib = load()
    n = len(shards)
    arr = (Shard * n)()
    unit = 2 if name.startswith("convert_utf16") else 1
    for i, (dev, t_in, t_out) in enumerate(shards):
        arr[i].device = dev
        arr[i].d_in = t_in.data_ptr() if t_in.numel() else None
        arr[i].len = (t_in.numel() * t_in.element_size()) // unit
        arr[i].d_out = t_out.data_ptr() if t_out is not None else None
    res = (ShardedResultC * n)()
    _check(getattr(lib, f"b200_mgpu_{name}")(arr, n, res), f"b200_mgpu_{name}")
    return [r.astuple() for r in res]


def launch_count() -> int:
    return int(load().b200_launch_count())


# ---------------------------------------------------------------------------------------------------
# Argument marshalling: CUDA tensors go to the device entry points, everything else to b200_host_*.
# ---------------------------------------------------------------------------------------------------
def _is_cuda_tensor(x) -> bool:
    return hasattr(x, "is_cuda") and bool(x.is_cuda)


def _stream_of(x) -> int:
    import torch
    return int(torch.cuda.current_stream(x.device).cuda_stream)


def _host_view(data, unit: int):
    """(address, length in units, keepalive) of a host buffer: bytes, bytearray, numpy array or CPU tensor."""
    import numpy as np
    if hasattr(data, "data_ptr") and not _is_cuda_tensor(data):  # CPU torch tensor (possibly pinned)
        if not data.is_contiguous():
            raise B200Error("host tensor must be contiguous")
        nbytes = data.numel() * data.element_size()
        return int(data.data_ptr()), nbytes // unit, data
    if isinstance(data, (bytes, bytearray, memoryview)):
        arr = np.frombuffer(data, dtype=np.uint8)
    else:
        arr = np.ascontiguousarray(data)
    nbytes = arr.size * arr.itemsize
    return (int(arr.ctypes.data) if nbytes else 0), nbytes // unit, arr


def _dev_view(t, unit: int):
    if not t.is_contiguous():
        raise B200Error("device tensor must be contiguous")
    return int(t.data_ptr()), (t.numel() * t.element_size()) // unit


def _reduce_op(name: str, data, unit: int, res):
    lib = load()
    if _is_cuda_tensor(data):
        ptr, n = _dev_view(data, unit)
        _check(getattr(lib, f"b200_{name}")(ptr, n, ctypes.byref(res), _stream_of(data)), name)
    else:
        ptr, n, _keep = _host_view(data, unit)
        _check(getattr(lib, f"b200_host_{name}")(ptr, n, ctypes.byref(res)), name)
    return res


def validate_utf8_with_errors(data):
    """simdutf::validate_utf8_with_errors -> (error, count)."""
    return _reduce_op("validate_utf8_with_errors", data, 1, Result()).astuple()


def validate_utf8(data) -> bool:
    return validate_utf8_with_errors(data)[0] == SUCCESS


def count_utf8(data) -> int:
    return int(_reduce_op("count_utf8", data, 1, ctypes.c_uint64()).value)


def utf32_length_from_utf8(data) -> int:
    return count_utf8(data)


def utf16_length_from_utf8(data) -> int:
    return int(_reduce_op("utf16_length_from_utf8", data, 1, ctypes.c_uint64()).value)


def count_utf16le(data) -> int:
    return int(_reduce_op("count_utf16le", data, 2, ctypes.c_uint64()).value)


def utf32_length_from_utf16le(data) -> int:
    return count_utf16le(data)


def utf8_length_from_utf16le(data) -> int:
    return int(_reduce_op("utf8_length_from_utf16le", data, 2, ctypes.c_uint64()).value)


def validate_utf16le_with_errors(data):
    return _reduce_op("validate_utf16le_with_errors", data, 2, Result()).astuple()


def _convert_op(name: str, data, in_unit: int, out):
    """out: CUDA tensor (device path) or writable numpy array / CPU tensor (host path), caller-sized."""
    lib = load()
    res = Result()
    if _is_cuda_tensor(data):
        ptr, n = _dev_view(data, in_unit)
        if not _is_cuda_tensor(out):
            raise B200Error("device input needs a device output buffer")
        _check(getattr(lib, f"b200_{name}")(ptr, n, int(out.data_ptr()), ctypes.byref(res), _stream_of(data)), name)
    else:
        ptr, n, _keep = _host_view(data, in_unit)
        optr, _on, _okeep = _host_view(out, 1)
        _check(getattr(lib, f"b200_host_{name}")(ptr, n, optr, ctypes.byref(res)), name)
    return res.astuple()


def convert_utf8_to_utf16le_with_errors(data, out):
    """simdutf::convert_utf8_to_utf16le_with_errors -> (error, units written | error position)."""
    return _convert_op("convert_utf8_to_utf16le", data, 1, out)


def convert_utf8_to_utf16le(data, out) -> int:
    """simdutf::convert_utf8_to_utf16le -> units written, 0 on any error."""
    err, count = convert_utf8_to_utf16le_with_errors(data, out)
    return 0 if err else count


def convert_utf8_to_utf32_with_errors(data, out):
    return _convert_op("convert_utf8_to_utf32", data, 1, out)


def convert_utf8_to_utf32(data, out) -> int:
    err, count = convert_utf8_to_utf32_with_errors(data, out)
    return 0 if err else count


def convert_utf16le_to_utf8_with_errors(data, out):
    return _convert_op("convert_utf16le_to_utf8", data, 2, out)


def convert_utf16le_to_utf8(data, out) -> int:
    err, count = convert_utf16le_to_utf8_with_errors(data, out)
    return 0 if err else count


# ---- UTF-16BE twins: `data` / `out` hold big-endian 16-bit units ------------------------------
def count_utf16be(data) -> int:
    return int(_reduce_op("count_utf16be", data, 2, ctypes.c_uint64()).value)


def utf32_length_from_utf16be(data) -> int:
    return count_utf16be(data)


def utf8_length_from_utf16be(data) -> int:
    return int(_reduce_op("utf8_length_from_utf16be", data, 2, ctypes.c_uint64()).value)


def validate_utf16be_with_errors(data):
    return _reduce_op("validate_utf16be_with_errors", data, 2, Result()).astuple()


def convert_utf8_to_utf16be_with_errors(data, out):
    return _convert_op("convert_utf8_to_utf16be", data, 1, out)


def convert_utf16be_to_utf8_with_errors(data, out):
    return _convert_op("convert_utf16be_to_utf8", data, 2, out)


def change_endianness_utf16(data, out) -> None:
    """simdutf::change_endianness_utf16: out[i] = byteswap(data[i])."""
    err, _n = _convert_op("change_endianness_utf16", data, 2, out)
    if err:
        raise B200Error("change_endianness_utf16 failed")


# ---- UTF-32 family (second part) ---------------------------------------------------------------
def validate_utf32_with_errors(data):
    return _reduce_op("validate_utf32_with_errors", data, 4, Result()).astuple()


def utf8_length_from_utf32(data) -> int:
    return int(_reduce_op("utf8_length_from_utf32", data, 4, ctypes.c_uint64()).value)


def utf16_length_from_utf32(data) -> int:
    return int(_reduce_op("utf16_length_from_utf32", data, 4, ctypes.c_uint64()).value)


def convert_utf32_to_utf8_with_errors(data, out):
    return _convert_op("convert_utf32_to_utf8", data, 4, out)


def convert_utf32_to_utf16le_with_errors(data, out):
    return _convert_op("convert_utf32_to_utf16le", data, 4, out)


def convert_utf32_to_utf16be_with_errors(data, out):
    return _convert_op("convert_utf32_to_utf16be", data, 4, out)


def convert_utf16le_to_utf32_with_errors(data, out):
    return _convert_op("convert_utf16le_to_utf32", data, 2, out)


def convert_utf16be_to_utf32_with_errors(data, out):
    return _convert_op("convert_utf16be_to_utf32", data, 2, out)


# ---- Latin-1 / ASCII family --------------------------------------------------------------------
def validate_ascii_with_errors(data):
    return _reduce_op("validate_ascii_with_errors", data, 1, Result()).astuple()


def utf8_length_from_latin1(data) -> int:
    return int(_reduce_op("utf8_length_from_latin1", data, 1, ctypes.c_uint64()).value)


def latin1_length_from_utf8(data) -> int:
    return count_utf8(data)


def convert_latin1_to_utf8(data, out):
    return _convert_op("convert_latin1_to_utf8", data, 1, out)


def convert_latin1_to_utf16le(data, out):
    return _convert_op("convert_latin1_to_utf16le", data, 1, out)


def convert_latin1_to_utf16be(data, out):
    return _convert_op("convert_latin1_to_utf16be", data, 1, out)


def convert_latin1_to_utf32(data, out):
    return _convert_op("convert_latin1_to_utf32", data, 1, out)


def convert_utf8_to_latin1_with_errors(data, out):
    return _convert_op("convert_utf8_to_latin1", data, 1, out)


def convert_utf16le_to_latin1_with_errors(data, out):
    return _convert_op("convert_utf16le_to_latin1", data, 2, out)


def convert_utf16be_to_latin1_with_errors(data, out):
    return _convert_op("convert_utf16be_to_latin1", data, 2, out)


def convert_utf32_to_latin1_with_errors(data, out):
    return _convert_op("convert_utf32_to_latin1", data, 4, out)


# ---- SURVEY.md §8f rank 4 ---------------------------------------------------------------------------------------------
def to_well_formed_utf16le(data, out) -> None:
    _convert_op("to_well_formed_utf16le", data, 2, out)


def to_well_formed_utf16be(data, out) -> None:
    _convert_op("to_well_formed_utf16be", data, 2, out)


def detect_encodings(data) -> int:
    """OR of encoding_type values: UTF8 = 1, UTF16_LE = 2, UTF16_BE = 4, UTF32_LE = 8, UTF32_BE = 16."""
    return int(_reduce_op("detect_encodings", data, 1, ctypes.c_uint64()).value)


def base64_length_from_binary(length: int, options: int = 0) -> int:
    return int(load().b200_base64_length_from_binary(length, options))


def binary_to_base64(data, out, options: int = 0) -> int:
    """simdutf::binary_to_base64 -> characters written (out must hold base64_length_from_binary(len, options))."""
    lib = load()
    res = Result()
    if _is_cuda_tensor(data):
        ptr, n = _dev_view(data, 1)
        if not _is_cuda_tensor(out):
            raise B200Error("device input needs a device output buffer")
        _check(lib.b200_binary_to_base64(ptr, n, int(out.data_ptr()), options, ctypes.byref(res), _stream_of(data)), "binary_to_base64")
    else:
        ptr, n, _keep = _host_view(data, 1)
        optr, _on, _okeep = _host_view(out, 1)
        _check(lib.b200_host_binary_to_base64(ptr, n, optr, options, ctypes.byref(res)), "binary_to_base64")
    return int(res.count)


def base64_to_binary_details(data, out, options: int = base64_default, last_chunk: int = loose):
    """simdutf::base64_to_binary_details -> (error, input_count, output_count)."""
    lib = load()
    res = FullResult()
    if _is_cuda_tensor(data):
        ptr, n = _dev_view(data, 1)
        _check(lib.b200_base64_to_binary(ptr, n, int(out.data_ptr()), options, last_chunk, ctypes.byref(res),
                                         _stream_of(data)), "base64_to_binary")
    else:
        ptr, n, _keep = _host_view(data, 1)
        optr, _on, _okeep = _host_view(out, 1)
        _check(lib.b200_host_base64_to_binary(ptr, n, optr, options, last_chunk, ctypes.byref(res)), "base64_to_binary")
    return res.astuple()


def base64_to_binary_details_utf16(data, out, options: int = base64_default, last_chunk: int = loose):
    """simdutf::base64_to_binary_details for char16_t input (`data`: 16-bit units) -> (error, input_count, output_count)."""
    lib = load()
    res = FullResult()
    if _is_cuda_tensor(data):
        ptr, n = _dev_view(data, 2)
        _check(lib.b200_base64_to_binary_utf16(ptr, n, int(out.data_ptr()), options, last_chunk, ctypes.byref(res),
                                               _stream_of(data)), "base64_to_binary_utf16")
    else:
        ptr, n, _keep = _host_view(data, 2)
        optr, _on, _okeep = _host_view(out, 1)
        _check(lib.b200_host_base64_to_binary_utf16(ptr, n, optr, options, last_chunk, ctypes.byref(res)), "base64_to_binary_utf16")
    return res.astuple()


def base64_to_binary(data, out, options: int = base64_default, last_chunk: int = loose):
    """simdutf::base64_to_binary -> (error, count) via the reference's full_result -> result rule."""
    e, i, o = base64_to_binary_details(data, out, options, last_chunk)
    return (e, o) if e in (SUCCESS, BASE64_INPUT_REMAINDER) else (e, i)


def maximal_binary_length_from_base64(data: bytes) -> int:
    ptr, n, _keep = _host_view(data, 1)
    return int(load().b200_host_maximal_binary_length_from_base64(ptr, n))


def trim_partial_utf8(data: bytes) -> int:
    ptr, n, _keep = _host_view(data, 1)
    return int(load().b200_host_trim_partial_utf8(ptr, n))


# ---------------------------------------------------------------------------------------------
# Many small strings per launch (include/simdutf_b200.h, "Many small strings per launch"; kernels in csrc/k_batch.cu)
# ---------------------------------------------------------------------------------------------
def _batch_views(strings):
    n = len(strings)
    bufs = [bytes(x) for x in strings]
    ptrs = (ctypes.c_char_p * max(n, 1))(*bufs)
    lens = (ctypes.c_size_t * max(n, 1))(*[len(x) for x in bufs])
    return n, bufs, ptrs, lens


def validate_utf8_batch(strings):
    """[(error, count)] of simdutf::validate_utf8_with_errors for every string of a list of host byte strings: one
    upload, one launch, one download."""
    n, _keep, ptrs, lens = _batch_views(strings)
    res = (Result * max(n, 1))()
    _check(load().b200_host_validate_utf8_batch(ptrs, lens, n, res), "validate_utf8_batch")
    return [res[i].astuple() for i in range(n)]


def utf16_length_from_utf8_batch(strings):
    n, _keep, ptrs, lens = _batch_views(strings)
    out = (ctypes.c_uint64 * max(n, 1))()
    _check(load().b200_host_utf16_length_from_utf8_batch(ptrs, lens, n, out), "utf16_length_from_utf8_batch")
    return [int(out[i]) for i in range(n)]


def count_utf8_batch(strings):
    n, _keep, ptrs, lens = _batch_views(strings)
    out = (ctypes.c_uint64 * max(n, 1))()
    _check(load().b200_host_count_utf8_batch(ptrs, lens, n, out), "count_utf8_batch")
    return [int(out[i]) for i in range(n)]


def convert_utf8_to_utf16le_batch(strings):
    """[((error, count), units-as-bytes or None)] for every string; a string's output buffer is sized by its byte
    length (never fewer units than that)."""
    n, bufs, ptrs, lens = _batch_views(strings)
    outs = [ctypes.create_string_buffer(2 * len(x) + 2) for x in bufs]
    optrs = (ctypes.c_void_p * max(n, 1))(*[ctypes.addressof(o) for o in outs])
    res = (Result * max(n, 1))()
    _check(load().b200_host_convert_utf8_to_utf16le_batch(ptrs, lens, n, optrs, res), "convert_utf8_to_utf16le_batch")
    return [(res[i].astuple(), outs[i].raw[: 2 * res[i].count] if res[i].error == 0 else None) for i in range(n)]


def validate_utf8_batch_device(data, offsets, results=None):
    """Device flavour: `data` uint8 CUDA tensor, `offsets` int64/uint64 CUDA tensor of n + 1 entries; returns an int64 CUDA
    tensor [n, 2] viewed from b200_result (error in the low 32 bits of column 0)."""
    import torch
    n = offsets.numel() - 1
    if results is None:
        results = torch.empty((max(n, 1), 2), dtype=torch.int64, device=data.device)
    _check(load().b200_validate_utf8_batch_async(int(data.data_ptr()), int(offsets.data_ptr()), n, int(results.data_ptr()),
                                                  _stream_of(data)), "validate_utf8_batch_async")
    return results[:n]
