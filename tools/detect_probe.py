"""Times the three validations detect_encodings is made of, on the same UTF-16LE text."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import simdutf_b200 as b
from simdutf_b200 import synth
lib = b.load(); b.set_device(0)
dev = torch.device("cuda", 0)
sp = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
res = torch.zeros(4, dtype=torch.int64, device=dev); rp = ctypes.c_void_p(res.data_ptr())
u = synth.mixed_utf16le(1 << 29, seed=3, device=dev)
nb = u.numel() * 2 // 4 * 4
p = ctypes.c_void_p(u.data_ptr())
def t(name, fn):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): fn()
    e1.record(); torch.cuda.synchronize()
    print(f"{name}: {e0.elapsed_time(e1) / 5:.4f} ms  result {res.tolist()}")
t("validate_utf8 on utf16 text", lambda: lib.b200_validate_utf8_with_errors_async(p, nb, rp, sp))
t("validate_utf16le", lambda: lib.b200_validate_utf16le_with_errors_async(p, nb // 2, rp, sp))
t("validate_utf32 on utf16 text", lambda: lib.b200_validate_utf32_with_errors_async(p, nb // 4, rp, sp))
t("detect_encodings", lambda: lib.b200_detect_encodings_async(p, nb, rp, sp))
