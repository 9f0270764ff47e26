"""Small run of every hot-path operation for compute-sanitizer (memcheck / racecheck / initcheck are slow: keep
the inputs at a few hundred KiB).  usage: compute-sanitizer --tool memcheck python tools/sanitize_smoke.py"""
import os
import random
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import simdutf_b200 as b
from simdutf_b200 import synth

b.load()
b.set_device(0)
dev = torch.device("cuda", 0)
rng = random.Random(1)


def place(t, mis):
    buf = torch.zeros(t.numel() + mis + 64, dtype=t.dtype, device=dev)
    buf[mis:mis + t.numel()] = t
    return buf[mis:mis + t.numel()]


for n, mis in ((300_000, 0), (70_001, 5), (2_049, 13), (17, 1)):
    d = place(synth.mixed_utf8(n, seed=2, device=dev), mis)
    assert b.validate_utf8_with_errors(d)[0] == 0
    units = b.utf16_length_from_utf8(d)
    o16 = place(torch.zeros(units, dtype=torch.int16, device=dev), mis % 8)
    assert b.convert_utf8_to_utf16le_with_errors(d, o16) == (0, units)
    c = b.count_utf8(d)
    o32 = place(torch.zeros(c, dtype=torch.int32, device=dev), mis % 4)
    assert b.convert_utf8_to_utf32_with_errors(d, o32) == (0, c)
    nb = b.utf8_length_from_utf16le(o16)
    assert nb == d.numel()
    o8 = place(torch.zeros(nb, dtype=torch.uint8, device=dev), mis)
    assert b.convert_utf16le_to_utf8_with_errors(o16, o8) == (0, nb)
    assert torch.equal(o8, d)
    bad = d.clone()
    bad[bad.numel() // 2] = 0xFF
    assert b.convert_utf8_to_utf16le_with_errors(bad, o16)[0] != 0
    text, payload = synth.base64_text(max(n, 64), seed=4, device=dev)
    t = place(text, mis)
    ob = place(torch.zeros(t.numel(), dtype=torch.uint8, device=dev), (mis * 7) % 16)
    e, i, k = b.base64_to_binary_details(t, ob, 0, 0)
    assert e == 0 and k == payload.numel() and torch.equal(ob[:k], payload)
a = place(synth.ascii_text(100_000, seed=1, device=dev), 3)
assert b.validate_utf8_with_errors(a) == (0, a.numel())
o = torch.zeros(a.numel(), dtype=torch.int16, device=dev)
assert b.convert_utf8_to_utf16le_with_errors(a, o) == (0, a.numel())
torch.cuda.synchronize()
print("sanitize_smoke ok")
