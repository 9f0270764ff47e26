"""Builds tests/golden/golden.json.  Run in the BUILD container (needs oracle/_ref/libsimdutf_ref.so, i.e.
the unmodified reference compiled from /root/reference by `make -C oracle ref`):

    python tests/golden/make_golden.py

Two kinds of vectors:
  kat      known-answer tests transcribed from the reference's own test files, with the answers those
           tests assert (source file:line recorded per vector).  They pin the oracle independently of any
           library run.
  recorded outputs of the reference library itself (kernel named in "impl": the best kernel of this host,
           icelake here) on a seeded set of small inputs for every hot-path function, so the same answers
           can be checked on the GPU box, where /root/reference does not exist.
"""
import json
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from tests._oracle import Reference  # noqa: E402

SUCCESS, HEADER_BITS, TOO_SHORT, TOO_LONG, OVERLONG, TOO_LARGE, SURROGATE, INVALID_B64, B64_REMAINDER, B64_EXTRA = range(10)

kat = []


def K(func, data, expect, src, **kw):
    kat.append(dict(func=func, input=bytes(data).hex(), expect=expect, source=src, **kw))


# --- validate_utf8 (bool) : tests/validate_utf8_basic_tests.cpp:21-110 ---------------------------------------
good = [b"a", b"\xc3\xb1", b"\xe2\x82\xa1", b"\xf0\x90\x8c\xbc", b"\xc2\x80", b"\xf0\x90\x80\x80", b"\xee\x80\x80", b"\xef\xbb\xbf"]
bad = [b"\xc3\x28", b"\xa0\xa1", b"\xe2\x28\xa1", b"\xe2\x82\x28", b"\xf0\x28\x8c\xbc", b"\xf0\x90\x28\xbc", b"\xf0\x28\x8c\x28",
       b"\xc0\x9f", b"\xf5\xff\xff\xff", b"\xed\xa0\x81", b"\xf8\x90\x80\x80\x80", b"123456789012345\xed", b"123456789012345\xf1",
       b"123456789012345\xc2", b"\xC2\x7F", b"\xce", b"\xce\xba\xe1", b"\xce\xba\xe1\xbd", b"\xce\xba\xe1\xbd\xb9\xcf",
       b"\xce\xba\xe1\xbd\xb9\xcf\x83\xce", b"\xce\xba\xe1\xbd\xb9\xcf\x83\xce\xbc\xce", b"\xdf", b"\xef\xbf", b"\x80",
       b"\x91\x85\x95\x9e", b"\x6c\x02\x8e\x18"]
for g in good:
    K("validate_utf8", g, True, "tests/validate_utf8_basic_tests.cpp:21-28")
for b in bad:
    K("validate_utf8", b, False, "tests/validate_utf8_basic_tests.cpp:29-56")
K("validate_utf8", b"\x80", False, "tests/validate_utf8_basic_tests.cpp:6-10")
K("validate_utf8", b"\xC2\xA9", True, "tests/validate_utf8_basic_tests.cpp:12-16")
# tests/validate_utf8_puzzler_tests.cpp:6-14 (64 bytes, a lone 0x80 at offset 30)
puz = bytearray(64); puz[13] = 0x1c; puz[30] = 0x80
K("validate_utf8", puz, False, "tests/validate_utf8_puzzler_tests.cpp:6-14")

# --- validate_utf8_with_errors : exact (error, position) ----------------------------------------------------
K("validate_utf8_with_errors", b"\x20" * 63 + b"\xff", [HEADER_BITS, 63], "tests/validate_utf8_with_errors_tests.cpp:8-28")
bad102 = (b"\x0a\x04\x00\x00\xdb\xa1\xdd\xa1\xf1\xa0\xb6\x95\xe4\xb5\x89\xe7\x8f\x95"
          b"\xe4\xa2\x83\xe7\x95\x89\xe7\x95\x91\xe7\x95\x89\x00\x01\x01\x1a\x20\x28"
          b"\x00\x00\x60\x00\x00\x23\x00\xf1\xa0\xb6\x95\xe4\xb5\x89\xe7\x8f\x95\xe4"
          b"\xa2\x83\xe7\x95\x89\xe7\x95\x91\xe7\x81\x00\x00\x01\x01\x1a\x20\x28\x00"
          b"\x00\x60\x00\x00\x23\x00\x2f\x00\x00\x00\x00\x07\x04\x75\xc2\xa0\x34\x2f"
          b"\x00\x00\x00\x00\x07\x04\x75\xc2\xa0\x33\x53\x2b")
K("validate_utf8_with_errors", bad102, [TOO_SHORT, 62], "tests/validate_utf8_puzzler_tests.cpp:16-31")
K("validate_utf8_with_errors", b"\xC2\xA9", [SUCCESS, 2], "tests/validate_utf8_with_errors_tests.cpp:38-43")
K("validate_utf8_with_errors", b"\x20" * 64 + b"\xa9", [TOO_LONG, 64], "tests/convert_utf8_to_utf16le_with_errors_tests.cpp:22-36")

# --- convert_utf8_to_utf16le_with_errors : (error, position) ------------------------------------------------
K("convert_utf8_to_utf16le_with_errors", b"\x20" * 64 + b"\xa9", [TOO_LONG, 64], "tests/convert_utf8_to_utf16le_with_errors_tests.cpp:22-45")
a8 = (b"\x20" * 13 + b"\xf2\xa8\xa4\x8b" + b"\x20" * 14 + b"\xf2\xa8\xa4\x8b" + b"\x20" * 4 + b"\xf2\xa8\xa4\x8b" + b"\x20" * 2 +
      b"\xf2\xa8\xa4\x8b" + b"\x20" * 11 + b"\xf2\xa8\xa4\xa8\xa4" + b"\x20" * 63)
K("convert_utf8_to_utf16le_with_errors", a8, [TOO_LONG, 64], "tests/convert_utf8_to_utf16le_with_errors_tests.cpp:47-73")
i448 = b"\xcd\xb8" + b"\x20" * 61 + b"\xff" + b"\x20" * 64
K("convert_utf8_to_utf16le_with_errors", i448, [HEADER_BITS, 63], "tests/convert_utf8_to_utf16le_with_errors_tests.cpp:99-117")
K("convert_utf8_to_utf16le_with_errors", b"\x84", [TOO_LONG, 0], "tests/convert_utf8_to_utf16le_with_errors_tests.cpp:119-133")

# --- base64 (result = full_result -> result, include/simdutf/error.h:66-73) -----------------------------------
K("base64_to_binary", b" Y\fW\tJ\njZ A=\r= ", [SUCCESS, 4], "tests/base64_tests.cpp:1336-1338,1411-1423", options=0, last_chunk=0, output=b"abcd".hex())
simple = [(b"Hello, World!", b"SGVsbG8sIFdvcmxkIQ=="), (b"GeeksforGeeks", b"R2Vla3Nmb3JHZWVrcw=="), (b"123456", b"MTIzNDU2"),
          (b"Base64 Encoding", b"QmFzZTY0IEVuY29kaW5n")]
for plain, enc in simple:
    K("base64_to_binary", enc, [SUCCESS, len(plain)], "tests/base64_tests.cpp:1340-1350", options=0, last_chunk=0, output=plain.hex())
    K("base64_to_binary", enc.rstrip(b"="), [SUCCESS, len(plain)], "tests/base64_tests.cpp:1352-1373", options=1, last_chunk=0, output=plain.hex())
K("base64_to_binary", "Zg ==".encode(), [INVALID_B64, 2], "tests/base64_tests.cpp:1298-1318", options=0, last_chunk=0)
K("base64_to_binary", b"\x53\x53", [SUCCESS, 1], "tests/base64_tests.cpp:1320-1333", options=0, last_chunk=0)

# --- recorded from the reference library -----------------------------------------------------------------------
ref = Reference()
impl = ref.best
rng = random.Random(20261018)
recorded = []
special = [0x20, 0x41, 0x7F, 0x80, 0x8F, 0x90, 0x9F, 0xA0, 0xBF, 0xC0, 0xC1, 0xC2, 0xDF, 0xE0, 0xE1, 0xEC, 0xED, 0xEE, 0xEF, 0xF0,
           0xF1, 0xF3, 0xF4, 0xF5, 0xF7, 0xF8, 0xFF]


def rand_text(n):
    cps = []
    for _ in range(n):
        c = rng.randrange(4)
        cps.append(chr([rng.randrange(0x20, 0x7f), rng.randrange(0xa0, 0x250), rng.randrange(0x4e00, 0xa000),
                        rng.randrange(0x1f300, 0x1f650)][c]))
    return "".join(cps)


for i in range(120):
    n = rng.choice([0, 1, 2, 3, 5, 15, 16, 17, 31, 33, 63, 64, 65, 100, 127, 128, 129, 200, 500])
    mode = i % 4
    if mode == 0:
        d = bytes(rng.choice(special) for _ in range(n))
    elif mode == 1:
        d = rand_text(n).encode()
    else:
        b = bytearray(rand_text(n).encode())
        for _ in range(1 if b else 0):
            b[rng.randrange(len(b))] = rng.choice(special)
        d = bytes(b) if mode == 2 else bytes(b[: max(0, len(b) - rng.randrange(4))])
    r16, o16 = ref.convert_utf8_to_utf16le_with_errors(impl, d)
    r32, o32 = ref.convert_utf8_to_utf32_with_errors(impl, d)
    recorded.append(dict(kind="utf8", input=d.hex(), validate=list(ref.validate_utf8_with_errors(impl, d)),
                         count_utf8=ref.count_utf8(impl, d), utf16_length=ref.utf16_length_from_utf8(impl, d),
                         to_utf16=list(r16), utf16_out=o16.tobytes().hex(), to_utf32=list(r32), utf32_out=o32.tobytes().hex()))
import numpy as np  # noqa: E402
for i in range(80):
    n = rng.choice([0, 1, 2, 7, 8, 9, 31, 32, 33, 64, 100, 257])
    units = []
    while len(units) < n:
        c = rng.randrange(6)
        if i % 3 == 0:
            units.append(rng.choice([0x41, 0x7f, 0x80, 0x7ff, 0x800, 0xd7ff, 0xd800, 0xdbff, 0xdc00, 0xdfff, 0xe000, 0xffff]))
        elif c == 0:
            units.append(rng.randrange(0x80))
        elif c == 1:
            units.append(rng.randrange(0x80, 0x800))
        elif c == 2:
            units.append(rng.choice([rng.randrange(0x800, 0xd800), rng.randrange(0xe000, 0x10000)]))
        elif c == 3:
            units += [rng.randrange(0xd800, 0xdc00), rng.randrange(0xdc00, 0xe000)]
        elif i % 3 == 2 and rng.random() < 0.2:
            units.append(rng.randrange(0xd800, 0xe000))
        else:
            units.append(0x20)
    a = np.array(units, dtype=np.uint16)
    r8, o8 = ref.convert_utf16le_to_utf8_with_errors(impl, a)
    recorded.append(dict(kind="utf16", input=a.tobytes().hex(), count_utf16le=ref.count_utf16le(impl, a),
                         utf8_length=ref.utf8_length_from_utf16le(impl, a), validate=list(ref.validate_utf16le_with_errors(impl, a)),
                         to_utf8=list(r8), utf8_out=o8.tobytes().hex()))
abc = b"ABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz0123456789+/-_"
ws = b" \t\n\r\x0c"
fixed = [b"", b" ", b"=", b"==", b"A", b"AA", b"AAA", b"AAAA", b"AA=", b"AA==", b"AAA=", b"AAA==", b"AAAA=", b"A=", b"QQ===", b"AA=A",
         b"AAAA*AAA", b"AAAA AAAA\r\n", b"AAAAA", b" A A = = ", b"AAA\x80", b"AA-A", b"AA+A", b"AA_A", b"AA/A", b"QUJD REVG\n", b"QQ= ="]
for i in range(140):
    if i < len(fixed):
        d = fixed[i]
    else:
        n = rng.choice([3, 4, 5, 15, 16, 17, 63, 64, 65, 66, 100, 130, 300])
        out = bytearray()
        for _ in range(n):
            x = rng.random()
            if x < 0.82:
                out.append(rng.choice(abc[:64] if i % 2 else abc[:62] + abc[64:]))
            elif x < 0.95:
                out.append(rng.choice(ws))
            elif i % 5 == 0:
                out.append(rng.choice(b"=*\x80\xff.\x0b"))
            else:
                out.append(rng.choice(abc[:62]))
        d = bytes(out) + rng.choice([b"", b"=", b"==", b" = = ", b"= ", b"=\n=", b" "])
    rows = []
    for opt in (0, 1, 4, 5, 8, 12):
        for lc in (0, 1, 2):
            r, o = ref.base64_to_binary_details(impl, d, opt, lc)
            rows.append(dict(options=opt, last_chunk=lc, result=list(r), output=o.tobytes().hex() if r[0] not in (7, 9) else None))
    recorded.append(dict(kind="base64", input=d.hex(), maxlen=ref.maximal_binary_length_from_base64(d), cases=rows))

out = dict(reference="WojciechMula/simdutf v7.0.0 (/root/reference)", impl=impl, kat=kat, recorded=recorded)
path = os.path.join(ROOT, "tests", "golden", "golden.json")
with open(path, "w") as f:
    json.dump(out, f, separators=(",", ":"))
print(path, os.path.getsize(path), "bytes;", len(kat), "kat,", len(recorded), "recorded; impl", impl)
