"""Scratch: every conv_variant of convert_utf8_to_utf16le must produce the bytes of variant 1 (the previous kernel),
on misaligned inputs/outputs, with ASCII runs, and with an injected error; then time each at 1 GiB."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import simdutf_b200 as b
from simdutf_b200 import synth

lib = b.load(); b.set_device(0)
dev = torch.device("cuda", 0)
variants = [int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "0,2,3,4,5").split(",")]
sp = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
res = torch.zeros(4, dtype=torch.int64, device=dev)
rp = ctypes.c_void_p(res.data_ptr())

def conv(d, off_in, off_out, variant, be=False):
    b.set_tuning("conv_variant", variant)
    x = d[off_in:]
    n = x.numel()
    units = b.utf16_length_from_utf8(x)
    o = torch.full((units + 64,), 0x5A5A, dtype=torch.int16, device=dev)
    ov = o[off_out:off_out + units]
    fn = lib.b200_convert_utf8_to_utf16be_async if be else lib.b200_convert_utf8_to_utf16le_async
    rc = fn(ctypes.c_void_p(x.data_ptr()), n, ctypes.c_void_p(ov.data_ptr()), rp, sp)
    assert rc == 0, rc
    torch.cuda.synchronize()
    return res.tolist()[:2], o

bad = 0
base = synth.mixed_utf8(48 << 20, seed=7, device=dev)
# ASCII runs: make some 16 KiB stretches pure ASCII
mix = base.clone()
for k in range(0, mix.numel() - (1 << 16), 1 << 18):
    mix[k + 4096:k + 4096 + 40000] = 0x41
# fix up: the splice may cut characters -> produce a valid buffer by converting through the reference kernel? simply keep: invalid input is fine for
# comparing variants on the (error, count) result; outputs are compared only when valid
for name, d in (("mixed", base), ("ascii_runs", mix)):
    for off_in in (0, 1, 3, 17):
        for off_out in (0, 1, 2, 5, 7):
            ref_r, ref_o = conv(d, off_in, off_out, 1)
            for v in variants:
                r, o = conv(d, off_in, off_out, v)
                ok = r == ref_r and (ref_r[0] != 0 or torch.equal(o, ref_o))
                if not ok:
                    bad += 1
                    neq = (o != ref_o).nonzero()
                    print("MISMATCH", name, off_in, off_out, "variant", v, r, ref_r, "first diff", neq[:3].flatten().tolist(), "n diff", neq.numel())
# BE twin and small sizes against variant 1
for n in (0, 1, 2, 3, 31, 32, 33, 63, 64, 65, 1000, 2047, 2048, 2049, 14336, 14337, 100000):
    d = synth.mixed_utf8(n, seed=n + 1, device=dev) if n else torch.empty(0, dtype=torch.uint8, device=dev)
    for be in (False, True):
        ref_r, ref_o = conv(d, 0, 1, 1, be)
        for v in variants:
            r, o = conv(d, 0, 1, v, be)
            if r != ref_r or not torch.equal(o, ref_o):
                bad += 1
                print("MISMATCH small", n, be, "variant", v, r, ref_r)
# injected errors
d = base[: 8 << 20].clone()
for pos in (5, 2047, 2048, 14335, 14336, 1 << 20, (8 << 20) - 2):
    e = d.clone(); e[pos] = 0xFF
    ref_r, _ = conv(e, 0, 0, 1)
    for v in variants:
        r, _ = conv(e, 0, 0, v)
        if r != ref_r:
            bad += 1
            print("MISMATCH err", pos, "variant", v, r, ref_r)
print("variant check:", "FAILED %d" % bad if bad else "ok")

# timing
nbytes = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 30
d = synth.mixed_utf8(nbytes, seed=2, device=dev)
n = d.numel(); units = b.utf16_length_from_utf8(d)
o = torch.empty(units, dtype=torch.int16, device=dev)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
stag = [int(x) for x in (sys.argv[3] if len(sys.argv) > 3 else "0").split(",")]
for v, st in [(1, 0)] + [(v, st) for v in variants for st in stag]:
    b.set_tuning("conv_variant", v)
    b.set_tuning("conv_stagger", st)
    fn = lambda: lib.b200_convert_utf8_to_utf16le_async(ctypes.c_void_p(d.data_ptr()), n, ctypes.c_void_p(o.data_ptr()), rp, sp)
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(10): fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"variant {v} stagger {st}: {ms:.4f} ms/launch  in+out {(n + 2 * units) / ms / 1e6:.1f} GB/s  frac {(n + 2 * units) / ms / 1e6 / 6535.7:.3f}  result {res.tolist()[:2]}")
