"""Scratch: per-CTA clock instrumentation of k_utf8_transcode_v3 (dbg_lo/dbg_hi tuning knobs carry a device pointer)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import simdutf_b200 as b
from simdutf_b200 import synth
lib = b.load(); b.set_device(0)
dev = torch.device("cuda", 0)
sp = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
res = torch.zeros(4, dtype=torch.int64, device=dev)
rp = ctypes.c_void_p(res.data_ptr())
nbytes = 1 << 30
d = synth.mixed_utf8(nbytes, seed=2, device=dev)
n = d.numel(); units = b.utf16_length_from_utf8(d)
o = torch.empty(units, dtype=torch.int16, device=dev)
dbg = torch.zeros(16 * 1024, dtype=torch.int64, device=dev)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for v in [int(x) for x in sys.argv[1].split(",")]:
    b.set_tuning("conv_variant", v)
    fn = lambda: lib.b200_convert_utf8_to_utf16le_async(ctypes.c_void_p(d.data_ptr()), n, ctypes.c_void_p(o.data_ptr()), rp, sp)
    b.set_tuning("dbg_lo", 0); b.set_tuning("dbg_hi", 0)
    for _ in range(3): fn()
    torch.cuda.synchronize()
    p = dbg.data_ptr()
    lo, hi = p & 0xFFFFFFFF, p >> 32
    b.set_tuning("dbg_lo", lo - (1 << 32) if lo >= (1 << 31) else lo); b.set_tuning("dbg_hi", hi)
    dbg.zero_()
    e0.record(); fn(); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    b.set_tuning("dbg_lo", 0); b.set_tuning("dbg_hi", 0)
    t = dbg.view(-1, 16).cpu().double()  # signed sums
    t = t[t[:, 0] > 0]
    ctas = t.shape[0]
    it = t[:, 0]
    us = lambda x: x / 1965.0  # cycles -> microseconds at 1965 MHz
    print(f"variant {v}: {ms:.3f} ms with instrumentation, {ctas} CTAs, tiles/CTA mean {it.mean():.1f}")
    print(f"  scan warp per tile [us]: wait totals {us((t[:,1]/it).mean()):.2f}  look-back+post {us((t[:,2]/it).mean()):.2f}  max look-back {us(t[:,3].max()):.1f}  polls/tile {(t[:,4]/it).mean():.2f}")
    print(f"  look-back [us, globaltimer]: starts {(t[:,7].double()/it).mean()/1e3:.2f} after own aggregate; latest of the 32 nearest predecessors published {(t[:,5].view(torch.int64).double() if False else t[:,5]/it).mean()/1e3:.2f} after mine; batch 0 seen complete {(t[:,6]/it).mean()/1e3:.2f} after that")
    wi = t[:, 8]
    print(f"  worker 0 per iteration [us]: wait ticket {us((t[:,9]/wi).mean()):.2f}  pass 1 {us((t[:,10]/wi).mean()):.2f}  wait goff {us((t[:,11]/wi).mean()):.2f}  copy-out {us((t[:,12]/wi).mean()):.2f}  pass 2 {us((t[:,13]/wi).mean()):.2f}")
