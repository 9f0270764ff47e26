"""Tiny driver for ncu: runs ONE hot-path operation a few times on a device-resident synthetic buffer.
usage: python tools/prof_one.py <op> [bytes] [reps]
   op: convert16 | convert32 | validate_ascii | validate_mixed | length | utf16to8 | base64 |
       utf32to8 | utf32to16 | utf32to16be | utf16to32 | validate32 | len8from32 | b64encode |
       l1to8 | l1to16 | l1to32 | u8tol1 | u16tol1 | u32tol1 | validate_ascii_op | len8froml1 | wellformed | validate16
"""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import simdutf_b200 as b
from simdutf_b200 import synth

op = sys.argv[1]
nbytes = int(sys.argv[2]) if len(sys.argv) > 2 else 256 << 20
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
lib = b.load()
b.set_device(0)
for kv in filter(None, os.environ.get("B200_BENCH_TUNE", "").split(",")):  # experiments: "conv_minb=4"
    k, v = kv.split("=")
    b.set_tuning(k, int(v))
dev = torch.device("cuda", 0)
sp = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
res = torch.zeros(4, dtype=torch.int64, device=dev)
rp = ctypes.c_void_p(res.data_ptr())
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)


def run(fn, nin, nout):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(f"{op}: {ms:.4f} ms/launch, input {nin / ms / 1e6:.1f} GB/s, in+out {(nin + nout) / ms / 1e6:.1f} GB/s, result {res.tolist()}")


if op in ("convert16", "convert32", "validate_mixed", "length"):
    d = synth.mixed_utf8(nbytes, seed=2, device=dev)
    n = d.numel()
    p = ctypes.c_void_p(d.data_ptr())
    if op == "convert16":
        units = b.utf16_length_from_utf8(d)
        o = torch.empty(units, dtype=torch.int16, device=dev)
        run(lambda: lib.b200_convert_utf8_to_utf16le_async(p, n, ctypes.c_void_p(o.data_ptr()), rp, sp), n, 2 * units)
    elif op == "convert32":
        c = b.count_utf8(d)
        o = torch.empty(c, dtype=torch.int32, device=dev)
        run(lambda: lib.b200_convert_utf8_to_utf32_async(p, n, ctypes.c_void_p(o.data_ptr()), rp, sp), n, 4 * c)
    elif op == "validate_mixed":
        run(lambda: lib.b200_validate_utf8_with_errors_async(p, n, rp, sp), n, 0)
    else:
        run(lambda: lib.b200_utf16_length_from_utf8_async(p, n, rp, sp), n, 0)
elif op == "validate_ascii":
    d = synth.ascii_text(nbytes, seed=1, device=dev)
    n = d.numel()
    run(lambda: lib.b200_validate_utf8_with_errors_async(ctypes.c_void_p(d.data_ptr()), n, rp, sp), n, 0)
elif op in ("wellformed", "validate16"):
    u = synth.mixed_utf16le(nbytes // 2, seed=3, device=dev)
    n = u.numel()
    if op == "wellformed":
        o = torch.empty_like(u)
        run(lambda: lib.b200_to_well_formed_utf16le_async(ctypes.c_void_p(u.data_ptr()), n, ctypes.c_void_p(o.data_ptr()), rp, sp), 2 * n, 2 * n)
        assert torch.equal(o, u)
    else:
        run(lambda: lib.b200_validate_utf16le_with_errors_async(ctypes.c_void_p(u.data_ptr()), n, rp, sp), 2 * n, 0)
elif op == "utf16to8":
    u = synth.mixed_utf16le(nbytes // 2, seed=3, device=dev)
    n = u.numel()
    nb = b.utf8_length_from_utf16le(u)
    o = torch.empty(nb, dtype=torch.uint8, device=dev)
    run(lambda: lib.b200_convert_utf16le_to_utf8_async(ctypes.c_void_p(u.data_ptr()), n, ctypes.c_void_p(o.data_ptr()), rp, sp), 2 * n, nb)
elif op == "base64":
    t, pay = synth.base64_text(nbytes, seed=4, device=dev)
    n = t.numel()
    o = torch.empty(n // 4 * 3 + 3, dtype=torch.uint8, device=dev)
    run(lambda: lib.b200_base64_to_binary_async(ctypes.c_void_p(t.data_ptr()), n, ctypes.c_void_p(o.data_ptr()), 0, 0, rp, sp), n, pay.numel())
elif op in ("utf32to8", "utf32to16", "utf32to16be", "utf16to32", "validate32", "len8from32"):
    d = synth.mixed_utf8(nbytes, seed=2, device=dev)
    cps = b.count_utf8(d)
    u32 = torch.empty(cps, dtype=torch.int32, device=dev)
    assert b.convert_utf8_to_utf32_with_errors(d, u32) == (0, cps)
    p32 = ctypes.c_void_p(u32.data_ptr())
    units = b.utf16_length_from_utf8(d)
    if op == "utf32to8":
        o = torch.empty(d.numel(), dtype=torch.uint8, device=dev)
        run(lambda: lib.b200_convert_utf32_to_utf8_async(p32, cps, ctypes.c_void_p(o.data_ptr()), rp, sp), 4 * cps, d.numel())
        assert torch.equal(o, d)
    elif op in ("utf32to16", "utf32to16be"):
        o = torch.empty(units, dtype=torch.int16, device=dev)
        fn = lib.b200_convert_utf32_to_utf16le_async if op == "utf32to16" else lib.b200_convert_utf32_to_utf16be_async
        run(lambda: fn(p32, cps, ctypes.c_void_p(o.data_ptr()), rp, sp), 4 * cps, 2 * units)
    elif op == "utf16to32":
        u16 = torch.empty(units, dtype=torch.int16, device=dev)
        assert b.convert_utf8_to_utf16le_with_errors(d, u16) == (0, units)
        o = torch.empty(cps, dtype=torch.int32, device=dev)
        run(lambda: lib.b200_convert_utf16le_to_utf32_async(ctypes.c_void_p(u16.data_ptr()), units, ctypes.c_void_p(o.data_ptr()), rp, sp), 2 * units, 4 * cps)
        assert torch.equal(o, u32)
    elif op == "validate32":
        run(lambda: lib.b200_validate_utf32_with_errors_async(p32, cps, rp, sp), 4 * cps, 0)
    else:
        run(lambda: lib.b200_utf8_length_from_utf32_async(p32, cps, rp, sp), 4 * cps, 0)
elif op in ("l1to8", "l1to16", "l1to32", "u8tol1", "u16tol1", "u32tol1", "validate_ascii_op", "len8froml1"):
    g = torch.Generator(device=dev).manual_seed(59)
    lat = torch.randint(0, 0x80, (nbytes,), dtype=torch.uint8, device=dev, generator=g)
    if op != "validate_ascii_op":
        lat = torch.where(torch.rand(nbytes, device=dev, generator=g) < 0.3, lat | 0x80, lat)
    n = nbytes
    pl = ctypes.c_void_p(lat.data_ptr())
    n8 = b.utf8_length_from_latin1(lat)
    if op == "validate_ascii_op":
        run(lambda: lib.b200_validate_ascii_with_errors_async(pl, n, rp, sp), n, 0)
    elif op == "len8froml1":
        run(lambda: lib.b200_utf8_length_from_latin1_async(pl, n, rp, sp), n, 0)
    elif op == "l1to8":
        o = torch.empty(n8, dtype=torch.uint8, device=dev)
        run(lambda: lib.b200_convert_latin1_to_utf8_async(pl, n, ctypes.c_void_p(o.data_ptr()), rp, sp), n, n8)
    elif op == "l1to16":
        o = torch.empty(n, dtype=torch.int16, device=dev)
        run(lambda: lib.b200_convert_latin1_to_utf16le_async(pl, n, ctypes.c_void_p(o.data_ptr()), rp, sp), n, 2 * n)
    elif op == "l1to32":
        o = torch.empty(n, dtype=torch.int32, device=dev)
        run(lambda: lib.b200_convert_latin1_to_utf32_async(pl, n, ctypes.c_void_p(o.data_ptr()), rp, sp), n, 4 * n)
    elif op == "u8tol1":
        u8 = torch.empty(n8, dtype=torch.uint8, device=dev)
        assert b.convert_latin1_to_utf8(lat, u8) == (0, n8)
        o = torch.empty(n, dtype=torch.uint8, device=dev)
        run(lambda: lib.b200_convert_utf8_to_latin1_async(ctypes.c_void_p(u8.data_ptr()), n8, ctypes.c_void_p(o.data_ptr()), rp, sp), n8, n)
        assert torch.equal(o, lat)
    elif op == "u16tol1":
        u16 = lat.to(torch.int16)
        o = torch.empty(n, dtype=torch.uint8, device=dev)
        run(lambda: lib.b200_convert_utf16le_to_latin1_async(ctypes.c_void_p(u16.data_ptr()), n, ctypes.c_void_p(o.data_ptr()), rp, sp), 2 * n, n)
    else:
        u32 = lat.to(torch.int32)
        o = torch.empty(n, dtype=torch.uint8, device=dev)
        run(lambda: lib.b200_convert_utf32_to_latin1_async(ctypes.c_void_p(u32.data_ptr()), n, ctypes.c_void_p(o.data_ptr()), rp, sp), 4 * n, n)
elif op == "b64encode":
    pay = torch.randint(0, 256, (nbytes,), dtype=torch.uint8, device=dev)
    o = torch.empty((nbytes + 2) // 3 * 4, dtype=torch.uint8, device=dev)
    run(lambda: lib.b200_binary_to_base64_async(ctypes.c_void_p(pay.data_ptr()), nbytes, ctypes.c_void_p(o.data_ptr()), 0, rp, sp), nbytes, o.numel())
else:
    raise SystemExit("unknown op " + op)
