"""Builds tests/golden/golden_next.json: outputs of the unmodified reference library (oracle/_ref, kernel named in
"impl") on seeded small inputs for every function of SURVEY.md §8f ranks 1-4 (UTF-16BE twins, UTF-32 family,
Latin-1 / ASCII family, to_well_formed_utf16, detect_encodings, binary_to_base64).  Run in the BUILD container:

    python tests/golden/make_golden_next.py

tests/test_oracle.py::test_recorded_next checks oracle/oracle.c against these vectors, so the oracle stays pinned
where /root/reference does not exist.
"""
import json
import os
import random
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from tests._oracle import Reference  # noqa: E402

ref = Reference()
impl = [i for i in ("icelake", "haswell", "fallback") if i in ref.impls()][0]
rng = random.Random(20261018)
vec = []


def rec(func, data, out, **kw):
    vec.append(dict(func=func, input=bytes(data).hex(), out=out, **kw))


def res(r):  # ((error, count), array) -> json
    (e, c), o = r
    return [e, c, o.tobytes().hex()]


def units(n, dirty):
    u = []
    while len(u) < n:
        c = rng.randrange(6)
        if c == 0: u.append(rng.randrange(0x80))
        elif c == 1: u.append(rng.randrange(0x80, 0x800))
        elif c == 2: u.append(rng.choice([rng.randrange(0x800, 0xd800), rng.randrange(0xe000, 0x10000)]))
        elif c == 3: u += [rng.randrange(0xd800, 0xdc00), rng.randrange(0xdc00, 0xe000)]
        elif dirty and rng.random() < 0.15: u.append(rng.randrange(0xd800, 0xe000))
        else: u.append(0x20)
    return np.array(u, dtype=np.uint16)


def cps(n, dirty):
    out = []
    for _ in range(n):
        c = rng.randrange(6)
        out.append(rng.randrange(0x80) if c == 0 else rng.randrange(0x80, 0x800) if c == 1 else
                   rng.choice([rng.randrange(0x800, 0xd800), rng.randrange(0xe000, 0x10000)]) if c == 2 else
                   rng.randrange(0x10000, 0x110000) if c == 3 else
                   rng.choice([0x7f, 0x80, 0x7ff, 0x800, 0xd7ff, 0xe000, 0xffff, 0x10000, 0x10ffff]) if c == 4 else
                   (rng.choice([0xd800, 0xdfff, 0x110000, 0xffffffff]) if dirty and rng.random() < 0.2 else 0x41))
    return np.array(out, dtype=np.uint32)


u8pool = [0x00, 0x41, 0x7f, 0x80, 0xbf, 0xc0, 0xc1, 0xc2, 0xc3, 0xc4, 0xdf, 0xe0, 0xef, 0xf0, 0xf7, 0xf8, 0xff]
for it in range(60):
    n = rng.choice([0, 1, 2, 7, 15, 16, 17, 33, 64, 65, 100])
    dirty = it % 2 == 1
    # UTF-16BE twins + UTF-16 -> UTF-32 + to_well_formed
    u = units(n, dirty)
    for be in (False, True):
        src = u.byteswap() if be else u
        tag = "be" if be else "le"
        if be:
            rec("count_utf16be", src.tobytes(), ref.count_utf16be(impl, src))
            rec("utf8_length_from_utf16be", src.tobytes(), ref.utf8_length_from_utf16be(impl, src))
            rec("validate_utf16be_with_errors", src.tobytes(), list(ref.validate_utf16be_with_errors(impl, src)))
            rec("convert_utf16be_to_utf8_with_errors", src.tobytes(), res(ref.convert_utf16be_to_utf8_with_errors(impl, src)))
        rec(f"convert_utf16{tag}_to_utf32_with_errors", src.tobytes(), res(ref.convert_utf16_to_utf32_with_errors(impl, src, be)))
        rec(f"to_well_formed_utf16{tag}", src.tobytes(), ref.to_well_formed_utf16(impl, src, be).tobytes().hex())
    # UTF-32 family
    a = cps(n, dirty)
    rec("validate_utf32_with_errors", a.tobytes(), list(ref.validate_utf32_with_errors(impl, a)))
    rec("utf8_length_from_utf32", a.tobytes(), ref.utf8_length_from_utf32(impl, a))
    rec("utf16_length_from_utf32", a.tobytes(), ref.utf16_length_from_utf32(impl, a))
    rec("convert_utf32_to_utf8_with_errors", a.tobytes(), res(ref.convert_utf32_to_utf8_with_errors(impl, a)))
    rec("convert_utf32_to_utf16le_with_errors", a.tobytes(), res(ref.convert_utf32_to_utf16_with_errors(impl, a, False)))
    rec("convert_utf32_to_utf16be_with_errors", a.tobytes(), res(ref.convert_utf32_to_utf16_with_errors(impl, a, True)))
    a32 = np.array([rng.randrange(0x100) if rng.random() < 0.95 or not dirty else rng.choice([0x100, 0xffff, 0x10ffff, 0xffffff41])
                    for _ in range(n)], dtype=np.uint32)
    rec("convert_utf32_to_latin1_with_errors", a32.tobytes(), res(ref.convert_utf32_to_latin1_with_errors(impl, a32)))
    a16 = a32.astype(np.uint16) if not dirty else np.array([min(int(x), 0xffff) for x in a32], dtype=np.uint16)
    for be in (False, True):
        src = a16.byteswap() if be else a16
        rec(f"convert_utf16{'be' if be else 'le'}_to_latin1_with_errors", src.tobytes(), res(ref.convert_utf16_to_latin1_with_errors(impl, src, be)))
    # Latin-1 / ASCII
    lat = bytes(rng.randrange(256) if rng.random() < 0.3 else rng.randrange(0x80) for _ in range(n))
    rec("validate_ascii_with_errors", lat, list(ref.validate_ascii_with_errors(impl, lat)))
    rec("utf8_length_from_latin1", lat, ref.utf8_length_from_latin1(impl, lat))
    rec("convert_latin1_to_utf8", lat, ref.convert_latin1_to_utf8(impl, lat).tobytes().hex())
    rec("convert_latin1_to_utf16le", lat, ref.convert_latin1_to_utf16(impl, lat, False).tobytes().hex())
    rec("convert_latin1_to_utf16be", lat, ref.convert_latin1_to_utf16(impl, lat, True).tobytes().hex())
    rec("convert_latin1_to_utf32", lat, ref.convert_latin1_to_utf32(impl, lat).tobytes().hex())
    t = bytearray("".join(chr(rng.randrange(0x100) if rng.random() < 0.4 else rng.randrange(0x80)) for _ in range(n)).encode())
    if dirty and t:
        t[rng.randrange(len(t))] = rng.choice(u8pool)
    rec("latin1_length_from_utf8", t, ref.latin1_length_from_utf8(impl, bytes(t)))
    rec("convert_utf8_to_latin1_with_errors", t, res(ref.convert_utf8_to_latin1_with_errors(impl, bytes(t))))
    # detect_encodings
    text = "".join(rng.choice("aé中😀 ") for _ in range(n))
    d = rng.choice([b"", b"", b"\xff\xfe", b"\xfe\xff", b"\xef\xbb\xbf", b"\xff\xfe\x00\x00", b"\x00\x00\xfe\xff"]) + \
        text.encode(rng.choice(["utf-8", "utf-16-le", "utf-32-le"]))
    if dirty and d:
        k = rng.randrange(len(d)); d = d[:k] + bytes([rng.randrange(256)]) + d[k + 1:]
    rec("detect_encodings", d, ref.detect_encodings(impl, d))

out = dict(impl=impl, generator="tests/golden/make_golden_next.py", recorded=vec)
path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_next.json")
json.dump(out, open(path, "w"), indent=0)
print(len(vec), "vectors ->", path, os.path.getsize(path), "bytes")
