"""base64_to_binary on BASELINE config 4's text (CRLF every 76 characters + sparse blanks) and on text without any
whitespace, the output compared with the payload; then validate_utf32 and the pieces of detect_encodings.
(The geometry sweep and the A/B against round 1's two launches that this tool ran are in k_base64.cu and DESIGN.md §4.)
usage: python tools/prof_b64.py [bytes] [reps]
"""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import simdutf_b200 as b
from simdutf_b200 import synth

nbytes = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 30
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
lib = b.load()
b.set_device(0)
dev = torch.device("cuda", 0)
sp = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
res = torch.zeros(4, dtype=torch.int64, device=dev)
rp = ctypes.c_void_p(res.data_ptr())
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)


def run(name, fn, nin, nout):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(f"{name}: {ms:.4f} ms/call, input {nin / ms / 1e6:.1f} GB/s, in+out {(nin + nout) / ms / 1e6:.1f} GB/s", flush=True)


for label, kw in (("crlf76", {}), ("dense", {"line": 1 << 24, "sparse_ws": 0.0})):
    t, pay = synth.base64_text(nbytes, seed=4, device=dev, **kw)
    n = t.numel()
    o = torch.empty(n // 4 * 3 + 3, dtype=torch.uint8, device=dev)
    o.zero_()
    run(f"base64 {label}", lambda: lib.b200_base64_to_binary_async(ctypes.c_void_p(t.data_ptr()), n, ctypes.c_void_p(o.data_ptr()), 0, 0, rp, sp), n, pay.numel())
    r = res.tolist()
    ok = r[0] == 0 and r[2] == pay.numel() and bool(torch.equal(o[:pay.numel()], pay))
    print(f"   result {r[:3]} payload {pay.numel()} output {'OK' if ok else 'MISMATCH'}", flush=True)
    del t, pay, o
    torch.cuda.empty_cache()

d = synth.mixed_utf8(nbytes, seed=2, device=dev)
cps = b.count_utf8(d)
u32 = torch.empty(cps, dtype=torch.int32, device=dev)
assert b.convert_utf8_to_utf32_with_errors(d, u32) == (0, cps)
p32 = ctypes.c_void_p(u32.data_ptr())
run("validate_utf32", lambda: lib.b200_validate_utf32_with_errors_async(p32, cps, rp, sp), 4 * cps, 0)
print("   result", res.tolist()[:2], "want", [0, cps])
run("utf8_length_from_utf32", lambda: lib.b200_utf8_length_from_utf32_async(p32, cps, rp, sp), 4 * cps, 0)
print("   result", res.tolist()[:1], "want", d.numel())
u32[cps // 2] = 0x110000
run("validate_utf32 (error in the middle)", lambda: lib.b200_validate_utf32_with_errors_async(p32, cps, rp, sp), 4 * cps, 0)
print("   result", res.tolist()[:2], "want", [5, cps // 2])  # TOO_LARGE
del d, u32
u = synth.mixed_utf16le(nbytes // 2, seed=3, device=dev)
nb = u.numel() * 2 // 4 * 4
p = ctypes.c_void_p(u.data_ptr())
run("validate_utf8 on utf16 text", lambda: lib.b200_validate_utf8_with_errors_async(p, nb, rp, sp), nb, 0)
run("validate_utf16le", lambda: lib.b200_validate_utf16le_with_errors_async(p, nb // 2, rp, sp), nb, 0)
run("validate_utf32 on utf16 text", lambda: lib.b200_validate_utf32_with_errors_async(p, nb // 4, rp, sp), nb, 0)
run("detect_encodings on utf16 text", lambda: lib.b200_detect_encodings_async(p, nb, rp, sp), nb, 0)
print("   result", res.tolist()[:1])
a = synth.ascii_text(nbytes, seed=1, device=dev)
run("detect_encodings on ascii text", lambda: lib.b200_detect_encodings_async(ctypes.c_void_p(a.data_ptr()), a.numel() // 4 * 4, rp, sp), a.numel(), 0)
print("   result", res.tolist()[:1])
