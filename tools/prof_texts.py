"""convert_utf8_to_utf16le / utf16le_to_utf8 throughput on non-adversarial texts (device-resident, 256 MiB):
pure ASCII, ASCII with 5 % Latin-1 letters, Cyrillic words separated by spaces, CJK, emoji only."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import simdutf_b200 as b

lib = b.load(); b.set_device(0); dev = torch.device("cuda", 0)
sp = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
res = torch.zeros(4, dtype=torch.int64, device=dev); rp = ctypes.c_void_p(res.data_ptr())
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
rng = np.random.default_rng(7)
N = 1 << 22  # characters per base block, tiled up to the target size


def text(kind):
    if kind == "ascii":
        cps = rng.integers(0x20, 0x7F, N)
    elif kind == "ascii+5%latin":
        cps = np.where(rng.random(N) < 0.05, rng.integers(0xC0, 0x180, N), rng.integers(0x20, 0x7F, N))
    elif kind == "ascii+0.5%latin":
        cps = np.where(rng.random(N) < 0.005, rng.integers(0xC0, 0x180, N), rng.integers(0x20, 0x7F, N))
    elif kind == "pure cyrillic":
        cps = rng.integers(0x410, 0x450, N)
    elif kind == "cyrillic words":
        cps = np.where(rng.random(N) < 0.15, 0x20, rng.integers(0x410, 0x450, N))
    elif kind == "cjk":
        cps = rng.integers(0x4E00, 0x9FFF, N)
    elif kind == "emoji":
        cps = rng.integers(0x1F300, 0x1F650, N)
    s = "".join(map(chr, cps.tolist())).encode()
    reps = max(1, (256 << 20) // len(s))
    return torch.from_numpy(np.frombuffer(s * reps, dtype=np.uint8).copy()).to(dev)


def timeit(fn, reps=5):
    for _ in range(2): fn()
    torch.cuda.synchronize(); e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for kind in ("ascii", "ascii+0.5%latin", "ascii+5%latin", "pure cyrillic", "cyrillic words", "cjk", "emoji"):
    d = text(kind); n = d.numel(); p = ctypes.c_void_p(d.data_ptr())
    units = b.utf16_length_from_utf8(d)
    o = torch.empty(units, dtype=torch.int16, device=dev); op = ctypes.c_void_p(o.data_ptr())
    ms = timeit(lambda: lib.b200_convert_utf8_to_utf16le_async(p, n, op, rp, sp))
    assert res.tolist()[:2] == [0, units]
    o8 = torch.empty(n, dtype=torch.uint8, device=dev); o8p = ctypes.c_void_p(o8.data_ptr())
    ms2 = timeit(lambda: lib.b200_convert_utf16le_to_utf8_async(op, units, o8p, rp, sp))
    assert res.tolist()[:2] == [0, n] and torch.equal(o8, d)
    msv = timeit(lambda: lib.b200_validate_utf8_with_errors_async(p, n, rp, sp))
    print(f"{kind:16s} utf8->utf16 {n / ms / 1e6:7.1f} GB/s in ({(n + 2 * units) / ms / 1e6:7.1f} in+out)   "
          f"utf16->utf8 {2 * units / ms2 / 1e6:7.1f} GB/s in ({(2 * units + n) / ms2 / 1e6:7.1f} in+out)   validate {n / msv / 1e6:7.1f} GB/s")
