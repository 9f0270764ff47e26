"""Host-path (e2e) timing of utf16_length_from_utf8 + convert_utf8_to_utf16le on pinned host buffers, plus the raw
pinned H2D / D2H copy bandwidth of the box for reference."""
import ctypes, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import simdutf_b200 as b
from simdutf_b200 import synth
lib = b.load(); b.set_device(0); dev = torch.device("cuda", 0)
d = synth.mixed_utf8(1 << 30, seed=2, device=dev); n = d.numel()
units = b.utf16_length_from_utf8(d)
h_in = torch.empty(n, dtype=torch.uint8, pin_memory=True); h_in.copy_(d)
h_out = torch.empty(units, dtype=torch.int16, pin_memory=True)
if len(sys.argv) > 1 and sys.argv[1] == "raw":
    o = torch.empty(units, dtype=torch.int16, device=dev)
    for name, fn, nb in (("H2D", lambda: d.copy_(h_in, non_blocking=True), n), ("D2H", lambda: h_out.copy_(o, non_blocking=True), 2 * units)):
        fn(); torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(5): fn()
        torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 5
        print(f"raw pinned {name}: {nb / dt / 1e9:.1f} GB/s")
hres, hcnt = b.Result(), ctypes.c_uint64()
ip, op = ctypes.c_void_p(h_in.data_ptr()), ctypes.c_void_p(h_out.data_ptr())
def step():
    assert lib.b200_host_utf16_length_from_utf8(ip, n, ctypes.byref(hcnt)) == 0
    t1 = time.perf_counter()
    assert lib.b200_host_convert_utf8_to_utf16le(ip, n, op, ctypes.byref(hres)) == 0
    return t1
step()
t0 = time.perf_counter(); tl = tc = 0.0
for _ in range(4):
    a = time.perf_counter(); t1 = step(); c = time.perf_counter()
    tl += t1 - a; tc += c - t1
dt = (time.perf_counter() - t0) / 4
print(f"seg={os.environ.get('B200_TUNE_SEG_MB', '32')} MB: {dt * 1e3:.2f} ms/step ({n / dt / 1e9:.2f} GB/s): length {tl / 4 * 1e3:.2f} ms, convert {tc / 4 * 1e3:.2f} ms")
