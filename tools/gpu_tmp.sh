timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "utf16 or golden or config3" 2>&1 | tail -2
python - <<'PY'
import ctypes, torch, sys
sys.path.insert(0, '.')
import simdutf_b200 as b
from simdutf_b200 import synth
lib = b.load(); b.set_device(0); dev = torch.device('cuda', 0)
sp = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
res = torch.zeros(4, dtype=torch.int64, device=dev); rp = ctypes.c_void_p(res.data_ptr())
u = synth.mixed_utf16le(1 << 30, seed=3, device=dev); n = u.numel(); p = ctypes.c_void_p(u.data_ptr())
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for name, fn in (("count_utf16le", lib.b200_count_utf16le_async), ("utf8_length_from_utf16le", lib.b200_utf8_length_from_utf16le_async)):
    for _ in range(3): fn(p, n, rp, sp)
    torch.cuda.synchronize(); e0.record()
    for _ in range(10): fn(p, n, rp, sp)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(name, f"{ms:.4f} ms, {2*n/ms/1e6:.1f} GB/s", res.tolist()[:1])
PY
