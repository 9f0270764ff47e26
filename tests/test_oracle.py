"""CPU tests (-m "not gpu"): pin the oracle (oracle/oracle.c) against
  1. the golden vectors (tests/golden/golden.json): known-answer tests transcribed from the reference's own
     tests, and outputs recorded from the reference library by tests/golden/make_golden.py;
     (tests/golden/golden_next.json / make_golden_next.py: the same for the SURVEY.md §8f families);
  2. the unmodified reference library (oracle/_ref, icelake + haswell + fallback kernels) on seeded random
     inputs, when it is present (build container; it also travels to the GPU box as a prebuilt .so).
"""
import json
import os
import random

import numpy as np
import pytest

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "golden.json")


@pytest.fixture(scope="module")
def golden():
    with open(GOLDEN) as f:
        return json.load(f)


def test_kat(oracle, golden):
    for v in golden["kat"]:
        d = bytes.fromhex(v["input"])
        f = v["func"]
        if f == "validate_utf8":
            assert (oracle.validate_utf8_with_errors(d)[0] == 0) == v["expect"], v
        elif f == "validate_utf8_with_errors":
            assert list(oracle.validate_utf8_with_errors(d)) == v["expect"], v
        elif f == "convert_utf8_to_utf16le_with_errors":
            assert list(oracle.convert_utf8_to_utf16le_with_errors(d)[0]) == v["expect"], v
        elif f == "base64_to_binary":
            (e, i, o), out = oracle.base64_to_binary_details(d, v["options"], v["last_chunk"])
            res = [e, o] if e in (0, 8) else [e, i]
            assert res == v["expect"], v
            if "output" in v:
                assert out.tobytes().hex() == v["output"], v
        else:
            raise AssertionError(f)


def test_recorded(oracle, golden):
    n = 0
    for v in golden["recorded"]:
        d = bytes.fromhex(v["input"])
        if v["kind"] == "utf8":
            assert list(oracle.validate_utf8_with_errors(d)) == v["validate"]
            assert oracle.count_utf8(d) == v["count_utf8"]
            assert oracle.utf16_length_from_utf8(d) == v["utf16_length"]
            r, o = oracle.convert_utf8_to_utf16le_with_errors(d)
            assert list(r) == v["to_utf16"] and o.tobytes().hex() == v["utf16_out"]
            r, o = oracle.convert_utf8_to_utf32_with_errors(d)
            assert list(r) == v["to_utf32"] and o.tobytes().hex() == v["utf32_out"]
        elif v["kind"] == "utf16":
            a = np.frombuffer(d, dtype=np.uint16)
            assert oracle.count_utf16le(a) == v["count_utf16le"]
            assert oracle.utf8_length_from_utf16le(a) == v["utf8_length"]
            assert list(oracle.validate_utf16le_with_errors(a)) == v["validate"]
            r, o = oracle.convert_utf16le_to_utf8_with_errors(a)
            assert list(r) == v["to_utf8"] and o.tobytes().hex() == v["utf8_out"]
        else:
            assert oracle.maximal_binary_length_from_base64(d) == v["maxlen"]
            for c in v["cases"]:
                r, o = oracle.base64_to_binary_details(d, c["options"], c["last_chunk"])
                assert list(r)[:2] == c["result"][:2], (d, c, r)
                if c["output"] is not None:  # output_count is not pinned on errors 7/9
                    assert list(r) == c["result"] and o.tobytes().hex() == c["output"], (d, c, r)
        n += 1
    assert n == len(golden["recorded"]) and n > 300


SPECIAL = [0x20, 0x41, 0x7F, 0x80, 0x8F, 0x90, 0x9F, 0xA0, 0xBF, 0xC0, 0xC1, 0xC2, 0xDF, 0xE0, 0xE1, 0xEC, 0xED, 0xEE, 0xEF, 0xF0,
           0xF1, 0xF3, 0xF4, 0xF5, 0xF7, 0xF8, 0xFF]


def _rand_utf8(rng, n):
    mode = rng.randrange(4)
    if mode == 0:
        return bytes(rng.choice(SPECIAL) for _ in range(n))
    if mode == 1:
        return bytes(rng.randrange(256) for _ in range(n))
    s = "".join(chr(rng.choice([rng.randrange(0x20, 0x7f), rng.randrange(0xa0, 0x250), rng.randrange(0x4e00, 0xa000),
                                rng.randrange(0x1f300, 0x1f650)])) for _ in range(n // 2 + 1)).encode()[: n + 3]
    b = bytearray(s)
    if mode == 3 and b:
        b[rng.randrange(len(b))] = rng.choice(SPECIAL)
    return bytes(b)


def test_recorded_next(oracle):
    """the oracle against outputs recorded from the reference library
    (tests/golden/golden_next.json, made by tests/golden/make_golden_next.py) — pins these restatements where
    /root/reference is absent."""
    with open(os.path.join(os.path.dirname(__file__), "golden", "golden_next.json")) as f:
        g = json.load(f)

    def u16(d):
        return np.frombuffer(d, dtype=np.uint16)

    def u32(d):
        return np.frombuffer(d, dtype=np.uint32)

    def res(r):
        (e, c), o = r
        return [e, c, o.tobytes().hex()]

    table = {
        "count_utf16be": lambda d: oracle.count_utf16be(u16(d)),
        "utf8_length_from_utf16be": lambda d: oracle.utf8_length_from_utf16be(u16(d)),
        "validate_utf16be_with_errors": lambda d: list(oracle.validate_utf16be_with_errors(u16(d))),
        "convert_utf16be_to_utf8_with_errors": lambda d: res(oracle.convert_utf16be_to_utf8_with_errors(u16(d))),
        "convert_utf16le_to_utf32_with_errors": lambda d: res(oracle.convert_utf16_to_utf32_with_errors(u16(d), False)),
        "convert_utf16be_to_utf32_with_errors": lambda d: res(oracle.convert_utf16_to_utf32_with_errors(u16(d), True)),
        "to_well_formed_utf16le": lambda d: oracle.to_well_formed_utf16(u16(d), False).tobytes().hex(),
        "to_well_formed_utf16be": lambda d: oracle.to_well_formed_utf16(u16(d), True).tobytes().hex(),
        "validate_utf32_with_errors": lambda d: list(oracle.validate_utf32_with_errors(u32(d))),
        "utf8_length_from_utf32": lambda d: oracle.utf8_length_from_utf32(u32(d)),
        "utf16_length_from_utf32": lambda d: oracle.utf16_length_from_utf32(u32(d)),
        "convert_utf32_to_utf8_with_errors": lambda d: res(oracle.convert_utf32_to_utf8_with_errors(u32(d))),
        "convert_utf32_to_utf16le_with_errors": lambda d: res(oracle.convert_utf32_to_utf16_with_errors(u32(d), False)),
        "convert_utf32_to_utf16be_with_errors": lambda d: res(oracle.convert_utf32_to_utf16_with_errors(u32(d), True)),
        "convert_utf32_to_latin1_with_errors": lambda d: res(oracle.convert_utf32_to_latin1_with_errors(u32(d))),
        "convert_utf16le_to_latin1_with_errors": lambda d: res(oracle.convert_utf16_to_latin1_with_errors(u16(d), False)),
        "convert_utf16be_to_latin1_with_errors": lambda d: res(oracle.convert_utf16_to_latin1_with_errors(u16(d), True)),
        "validate_ascii_with_errors": lambda d: list(oracle.validate_ascii_with_errors(d)),
        "utf8_length_from_latin1": lambda d: oracle.utf8_length_from_latin1(d),
        "latin1_length_from_utf8": lambda d: oracle.count_utf8(d),
        "convert_latin1_to_utf8": lambda d: oracle.convert_latin1_to_utf8(d).tobytes().hex(),
        "convert_latin1_to_utf16le": lambda d: oracle.convert_latin1_to_utf16(d, False).tobytes().hex(),
        "convert_latin1_to_utf16be": lambda d: oracle.convert_latin1_to_utf16(d, True).tobytes().hex(),
        "convert_latin1_to_utf32": lambda d: oracle.convert_latin1_to_utf32(d).tobytes().hex(),
        "convert_utf8_to_latin1_with_errors": lambda d: res(oracle.convert_utf8_to_latin1_with_errors(d)),
        "detect_encodings": lambda d: oracle.detect_encodings(d),
    }
    seen = set()
    for v in g["recorded"]:
        d = bytes.fromhex(v["input"])
        assert table[v["func"]](d) == v["out"], v
        seen.add(v["func"])
    assert seen == set(table), set(table) - seen


def test_against_reference_library(oracle, ref):
    if ref is None:
        pytest.skip("oracle/_ref/libsimdutf_ref.so not built (no /root/reference here)")
    impls = [i for i in ("icelake", "haswell", "fallback") if i in ref.impls()]
    assert impls
    rng = random.Random(7)
    for _ in range(3000):
        d = _rand_utf8(rng, rng.randrange(0, 200))
        want = oracle.validate_utf8_with_errors(d)
        w16 = oracle.convert_utf8_to_utf16le_with_errors(d)
        w32 = oracle.convert_utf8_to_utf32_with_errors(d)
        for im in impls:
            assert ref.validate_utf8_with_errors(im, d) == want, (im, d.hex())
            assert ref.count_utf8(im, d) == oracle.count_utf8(d)
            assert ref.utf16_length_from_utf8(im, d) == oracle.utf16_length_from_utf8(d)
            r, o = ref.convert_utf8_to_utf16le_with_errors(im, d)
            assert r == w16[0] and o.tobytes() == w16[1].tobytes(), (im, d.hex())
            r, o = ref.convert_utf8_to_utf32_with_errors(im, d)
            assert r == w32[0] and o.tobytes() == w32[1].tobytes(), (im, d.hex())
        w16be = oracle.convert_utf8_to_utf16be_with_errors(d)
        assert w16be[0] == w16[0] and w16be[1].tobytes() == w16[1].byteswap().tobytes()
        if ref.has_be():
            for im in impls:
                r, o = ref.convert_utf8_to_utf16be_with_errors(im, d)
                assert r == w16be[0] and o.tobytes() == w16be[1].tobytes(), (im, d.hex())
    for _ in range(2000):
        n = rng.randrange(0, 80)
        u = []
        for _k in range(n):
            c = rng.randrange(6)
            if c == 0: u.append(rng.randrange(0x80))
            elif c == 1: u.append(rng.randrange(0x80, 0x800))
            elif c == 2: u.append(rng.choice([rng.randrange(0x800, 0xd800), rng.randrange(0xe000, 0x10000)]))
            elif c == 3: u += [rng.randrange(0xd800, 0xdc00), rng.randrange(0xdc00, 0xe000)]
            elif rng.random() < 0.1: u.append(rng.randrange(0xd800, 0xe000))
            else: u.append(0x20)
        a = np.array(u, dtype=np.uint16)
        want = oracle.convert_utf16le_to_utf8_with_errors(a)
        for im in impls:
            assert ref.count_utf16le(im, a) == oracle.count_utf16le(a)
            assert ref.utf8_length_from_utf16le(im, a) == oracle.utf8_length_from_utf16le(a)
            assert ref.validate_utf16le_with_errors(im, a) == oracle.validate_utf16le_with_errors(a)
            r, o = ref.convert_utf16le_to_utf8_with_errors(im, a)
            assert r == want[0] and o.tobytes() == want[1].tobytes(), (im, u)
        # the UTF-16BE twins on the byte-swapped buffer: same answers, pinned against the reference as well
        be = a.byteswap()
        assert oracle.count_utf16be(be) == oracle.count_utf16le(a)
        assert oracle.utf8_length_from_utf16be(be) == oracle.utf8_length_from_utf16le(a)
        assert oracle.validate_utf16be_with_errors(be) == oracle.validate_utf16le_with_errors(a)
        wbe = oracle.convert_utf16be_to_utf8_with_errors(be)
        assert wbe[0] == want[0] and wbe[1].tobytes() == want[1].tobytes()
        assert oracle.change_endianness_utf16(a).tobytes() == be.tobytes()
        if ref.has_be():
            for im in impls:
                assert ref.count_utf16be(im, be) == oracle.count_utf16be(be)
                assert ref.utf8_length_from_utf16be(im, be) == oracle.utf8_length_from_utf16be(be)
                assert ref.validate_utf16be_with_errors(im, be) == oracle.validate_utf16be_with_errors(be)
                r, o = ref.convert_utf16be_to_utf8_with_errors(im, be)
                assert r == wbe[0] and o.tobytes() == wbe[1].tobytes(), (im, u)
                assert ref.change_endianness_utf16(im, a).tobytes() == be.tobytes()
    # UTF-32 family: code points incl. the edges, sprinkled surrogates and out-of-range values
    for _ in range(1500):
        n = rng.randrange(0, 70)
        cps = []
        for _k in range(n):
            c = rng.randrange(8)
            cps.append(rng.randrange(0x80) if c == 0 else rng.randrange(0x80, 0x800) if c == 1 else
                       rng.choice([rng.randrange(0x800, 0xd800), rng.randrange(0xe000, 0x10000)]) if c == 2 else
                       rng.randrange(0x10000, 0x110000) if c == 3 else
                       rng.choice([0x7f, 0x80, 0x7ff, 0x800, 0xd7ff, 0xe000, 0xffff, 0x10000, 0x10ffff]) if c == 4 else
                       rng.choice([0xd800, 0xdbff, 0xdc00, 0xdfff, 0x110000, 0xffffffff]) if c == 5 and rng.random() < 0.08 else 0x41)
        a32 = np.array(cps, dtype=np.uint32)
        w8 = oracle.convert_utf32_to_utf8_with_errors(a32)
        assert oracle.validate_utf32_with_errors(a32)[0] == w8[0][0]
        if ref.has_utf32():
            for im in impls:
                assert ref.validate_utf32_with_errors(im, a32) == oracle.validate_utf32_with_errors(a32), (im, cps)
                assert ref.utf8_length_from_utf32(im, a32) == oracle.utf8_length_from_utf32(a32)
                assert ref.utf16_length_from_utf32(im, a32) == oracle.utf16_length_from_utf32(a32)
                r, o = ref.convert_utf32_to_utf8_with_errors(im, a32)
                assert r == w8[0] and o.tobytes() == w8[1].tobytes(), (im, cps)
                for be in (False, True):
                    want = oracle.convert_utf32_to_utf16_with_errors(a32, be)
                    r, o = ref.convert_utf32_to_utf16_with_errors(im, a32, be)
                    assert r == want[0] and o.tobytes() == want[1].tobytes(), (im, be, cps)
        if w8[0][0] == 0:  # and back through UTF-16 -> UTF-32
            for be in (False, True):
                u16 = oracle.convert_utf32_to_utf16_with_errors(a32, be)[1]
                back = oracle.convert_utf16_to_utf32_with_errors(u16, be)
                assert back[0] == (0, a32.size) and back[1].tobytes() == a32.tobytes()
    if ref.has_utf32():
        for _ in range(1500):
            a = np.array([rng.choice([0x41, 0x7ff, 0x4e2d, 0xd800 + rng.randrange(0x400), 0xdc00 + rng.randrange(0x400), 0xffff])
                          for _k in range(rng.randrange(0, 40))], dtype=np.uint16)
            for be in (False, True):
                src = a.byteswap() if be else a
                want = oracle.convert_utf16_to_utf32_with_errors(src, be)
                for im in impls:
                    r, o = ref.convert_utf16_to_utf32_with_errors(im, src, be)
                    assert r == want[0] and o.tobytes() == want[1].tobytes(), (im, be, a)
    # Latin-1 / ASCII family
    u8pool = [0x00, 0x41, 0x7f, 0x80, 0xbf, 0xc0, 0xc1, 0xc2, 0xc3, 0xc4, 0xdf, 0xe0, 0xef, 0xf0, 0xf7, 0xf8, 0xff]
    for it in range(3000):
        n = rng.randrange(0, 80)
        mode = it % 4
        if mode == 0:    # Latin-1 bytes
            d = bytes(rng.randrange(256) if rng.random() < 0.3 else rng.randrange(0x80) for _ in range(n))
        elif mode == 1:  # valid Latin-1-range UTF-8
            d = "".join(chr(rng.randrange(0x100) if rng.random() < 0.4 else rng.randrange(0x80)) for _ in range(n)).encode()
        elif mode == 2:  # the same with one byte damaged
            d = bytearray("".join(chr(rng.randrange(0x100) if rng.random() < 0.4 else rng.randrange(0x80)) for _ in range(n)).encode())
            if d:
                d[rng.randrange(len(d))] = rng.choice(u8pool)
            d = bytes(d)
        else:            # byte soup from the class edges
            d = bytes(rng.choice(u8pool) if rng.random() < 0.5 else 0x41 for _ in range(n))
        w1 = oracle.convert_utf8_to_latin1_with_errors(d)
        if ref.has_latin1():
            for im in impls:
                assert ref.validate_ascii_with_errors(im, d) == oracle.validate_ascii_with_errors(d), (im, d)
                assert ref.utf8_length_from_latin1(im, d) == oracle.utf8_length_from_latin1(d)
                assert ref.latin1_length_from_utf8(im, d) == oracle.count_utf8(d)
                assert ref.convert_latin1_to_utf8(im, d).tobytes() == oracle.convert_latin1_to_utf8(d).tobytes()
                assert ref.convert_latin1_to_utf32(im, d).tobytes() == oracle.convert_latin1_to_utf32(d).tobytes()
                for be in (False, True):
                    assert ref.convert_latin1_to_utf16(im, d, be).tobytes() == oracle.convert_latin1_to_utf16(d, be).tobytes()
                r, o = ref.convert_utf8_to_latin1_with_errors(im, d)
                assert r == w1[0] and o.tobytes() == w1[1].tobytes(), (im, d, r, w1[0])
        if w1[0][0] == 0:
            assert oracle.convert_latin1_to_utf8(w1[1]).tobytes() == d
            assert w1[0][1] == oracle.count_utf8(d)
    for it in range(1500):
        n = rng.randrange(0, 50)
        a16 = np.array([rng.randrange(0x100) if rng.random() < 0.97 else rng.choice([0x100, 0x7ff, 0xd800, 0xffff]) for _ in range(n)], dtype=np.uint16)
        a32 = np.array([rng.randrange(0x100) if rng.random() < 0.97 else rng.choice([0x100, 0xffff, 0x10ffff, 0xffffff41]) for _ in range(n)], dtype=np.uint32)
        if ref.has_latin1():
            for im in impls:
                for be in (False, True):
                    src = a16.byteswap() if be else a16
                    want = oracle.convert_utf16_to_latin1_with_errors(src, be)
                    r, o = ref.convert_utf16_to_latin1_with_errors(im, src, be)
                    assert r == want[0] and o.tobytes() == want[1].tobytes(), (im, be, a16)
                want = oracle.convert_utf32_to_latin1_with_errors(a32)
                r, o = ref.convert_utf32_to_latin1_with_errors(im, a32)
                assert r == want[0] and o.tobytes() == want[1].tobytes(), (im, a32)
    # to_well_formed_utf16 and detect_encodings
    if ref.has_rank4():
        for it in range(2000):
            a = np.array([rng.choice([0x41, 0x4e2d, 0xd800 + rng.randrange(0x400), 0xdc00 + rng.randrange(0x400), 0xfffd])
                          for _k in range(rng.randrange(0, 40))], dtype=np.uint16)
            for be in (False, True):
                src = a.byteswap() if be else a
                want = oracle.to_well_formed_utf16(src, be)
                for im in impls:
                    assert ref.to_well_formed_utf16(im, src, be).tobytes() == want.tobytes(), (im, be, a)
        boms = [b"\xff\xfe", b"\xff\xfe\x00\x00", b"\xfe\xff", b"\x00\x00\xfe\xff", b"\xef\xbb\xbf", b"\xef\xbb"]
        for it in range(3000):
            n = rng.randrange(0, 64)
            mode = it % 5
            if mode == 0:
                d = "".join(rng.choice("aé中😀 ") for _ in range(n)).encode("utf-8")
            elif mode == 1:
                d = "".join(rng.choice("aé中😀 ") for _ in range(n)).encode("utf-16-le")
            elif mode == 2:
                d = "".join(rng.choice("aé中😀 ") for _ in range(n)).encode("utf-32-le")
            elif mode == 3:
                d = bytes(rng.randrange(256) for _ in range(n))
            else:
                d = rng.choice(boms) + "".join(rng.choice("aé中") for _ in range(n)).encode(rng.choice(["utf-8", "utf-16-le", "utf-32-le"]))
            if rng.random() < 0.2 and d:
                k = rng.randrange(len(d))
                d = d[:k] + bytes([rng.randrange(256)]) + d[k + 1:]
            want = oracle.detect_encodings(d)
            for im in impls:
                assert ref.detect_encodings(im, d) == want, (im, d, want)
    abc = b"ABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz0123456789+/-_"
    simd = [i for i in impls if i != "fallback"] or impls
    for it in range(1500):
        n = rng.randrange(0, 150)
        out = bytearray()
        for _k in range(n):
            x = rng.random()
            if x < 0.8: out.append(rng.choice(abc[:64] if it % 2 else abc[:62] + abc[64:]))
            elif x < 0.93: out.append(rng.choice(b" \t\n\r\x0c"))
            elif x < 0.97 and it % 4 >= 2: out.append(rng.choice(b"=*\x80\xff\x00." + abc[62:]))
            else: out.append(rng.choice(abc[:62]))
        d = bytes(out) + rng.choice([b"", b"=", b"==", b" = = ", b"= ", b"=\n=", b"===", b" "])
        assert ref.maximal_binary_length_from_base64(d) == oracle.maximal_binary_length_from_base64(d)
        for opt in (0, 1, 2, 3, 4, 5, 8, 12):
            for lc in (0, 1, 2):
                want, wout = oracle.base64_to_binary_details(d, opt, lc)
                # the SIMD kernels are the parity target; in accept_garbage mode the fallback kernel reports a
                # different input_count (it strips trailing whitespace first) — a known intra-reference difference
                for im in (impls if opt not in (4, 5, 12) else simd):
                    r, o = ref.base64_to_binary_details(im, d, opt, lc)
                    assert r[:2] == want[:2], (im, opt, lc, d, r, want)
                    if r[0] not in (7, 9):
                        assert r == want and o.tobytes() == wout.tobytes(), (im, opt, lc, d, r, want)
