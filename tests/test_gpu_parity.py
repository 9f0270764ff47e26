"""GPU parity tests (-m gpu): the sm_100a kernels, called through the C ABI (include/simdutf_b200.h), against
the oracle (oracle/oracle.c — pinned by tests/test_oracle.py) and the committed golden vectors.

Bit-exact everywhere: validity flag, error code, error position, counts and output bytes.
Nothing here reads /root/reference; oracle/_ref (prebuilt reference library) is used only if it travelled.
"""
import ctypes
import json
import os
import random

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "golden.json")
GUARD = 64  # canary elements after every output buffer


@pytest.fixture(scope="module")
def b():
    import simdutf_b200
    simdutf_b200.load()
    assert simdutf_b200.device_count() >= 1, "no sm_100 device visible: the CUDA path cannot run"
    simdutf_b200.set_device(0)
    return simdutf_b200


@pytest.fixture(scope="module")
def golden():
    with open(GOLDEN) as f:
        return json.load(f)


def dev(data, dtype=torch.uint8, misalign=0):
    """Device copy of `data` (bytes / numpy) placed `misalign` elements past an aligned allocation."""
    a = np.frombuffer(bytes(data), dtype=np.uint8) if isinstance(data, (bytes, bytearray)) else np.ascontiguousarray(data)
    t = torch.from_numpy(a.copy()).view(dtype) if a.size else torch.empty(0, dtype=dtype)
    buf = torch.zeros(t.numel() + misalign + 16, dtype=dtype, device="cuda")
    buf[misalign:misalign + t.numel()] = t.cuda()
    return buf[misalign:misalign + t.numel()]


FILL = {torch.uint8: 0x5A, torch.int16: 0x5A5A, torch.int32: 0x5A5A5A5A}


def out_buf(n, dtype, misalign=0):
    buf = torch.full((n + misalign + GUARD,), FILL[dtype], dtype=dtype, device="cuda")
    return buf, buf[misalign:misalign + n + GUARD]


def check_guard(view, n, fill=None):
    tail = view[n:n + GUARD].cpu()
    assert bool((tail == FILL[view.dtype]).all()), "output buffer overrun"


# ------------------------------------------------------------------------------------------------------------
# helpers that run one input through both call flavours and compare with the oracle
# ------------------------------------------------------------------------------------------------------------
def run_utf8(b, oracle, data: bytes, misalign=0, host_too=True):
    want_v = oracle.validate_utf8_with_errors(data)
    want16, o16 = oracle.convert_utf8_to_utf16le_with_errors(data)
    want32, o32 = oracle.convert_utf8_to_utf32_with_errors(data)
    n16 = oracle.utf16_length_from_utf8(data)
    n8 = oracle.count_utf8(data)
    d = dev(data, misalign=misalign)
    assert b.validate_utf8_with_errors(d) == want_v, (data[:64].hex(), len(data), misalign)
    assert b.count_utf8(d) == n8
    assert b.utf16_length_from_utf8(d) == n16
    # output buffers sized EXACTLY by the length query, as the reference's callers do
    _, v16 = out_buf(n16, torch.int16, misalign=misalign % 8)
    assert b.convert_utf8_to_utf16le_with_errors(d, v16) == want16, (data[:64].hex(), len(data), misalign)
    check_guard(v16, n16, 0x5A5A)
    if want16[0] == 0:
        assert v16[:n16].cpu().numpy().view(np.uint16).tobytes() == o16.tobytes()
    _, v32 = out_buf(n8, torch.int32, misalign=misalign % 4)
    assert b.convert_utf8_to_utf32_with_errors(d, v32) == want32
    check_guard(v32, n8, 0x5A5A5A5A)
    if want32[0] == 0:
        assert v32[:n8].cpu().numpy().view(np.uint32).tobytes() == o32.tobytes()
    if host_too:
        assert b.validate_utf8_with_errors(data) == want_v
        assert b.count_utf8(data) == n8 and b.utf16_length_from_utf8(data) == n16
        h16 = np.full(n16 + GUARD, 0x5A5A, dtype=np.uint16)
        assert b.convert_utf8_to_utf16le_with_errors(data, h16) == want16
        assert (h16[n16:] == 0x5A5A).all()
        if want16[0] == 0:
            assert h16[:n16].tobytes() == o16.tobytes()
        h32 = np.full(n8 + GUARD, 0x5A5A5A5A, dtype=np.uint32)
        assert b.convert_utf8_to_utf32_with_errors(data, h32) == want32
        if want32[0] == 0:
            assert h32[:n8].tobytes() == o32.tobytes()


def run_utf16(b, oracle, units: np.ndarray, misalign=0, host_too=True):
    units = np.ascontiguousarray(units, dtype=np.uint16)
    want, o8 = oracle.convert_utf16le_to_utf8_with_errors(units)
    n8 = oracle.utf8_length_from_utf16le(units)
    d = dev(units.view(np.uint8), dtype=torch.uint8, misalign=2 * misalign).view(torch.int16)
    assert b.count_utf16le(d) == oracle.count_utf16le(units)
    assert b.utf8_length_from_utf16le(d) == n8
    assert b.validate_utf16le_with_errors(d) == oracle.validate_utf16le_with_errors(units)
    _, v8 = out_buf(n8, torch.uint8, misalign=misalign)
    assert b.convert_utf16le_to_utf8_with_errors(d, v8) == want, (units[:32], len(units), misalign)
    check_guard(v8, n8)
    if want[0] == 0:
        assert v8[:n8].cpu().numpy().tobytes() == o8.tobytes()
    if host_too:
        h8 = np.full(n8 + GUARD, 0x5A, dtype=np.uint8)
        assert b.convert_utf16le_to_utf8_with_errors(units, h8) == want
        assert (h8[n8:] == 0x5A).all()
        if want[0] == 0:
            assert h8[:n8].tobytes() == o8.tobytes()
        assert b.count_utf16le(units) == oracle.count_utf16le(units)


def run_b64(b, oracle, data: bytes, options=0, last_chunk=0, misalign=0, host_too=True):
    want, wout = oracle.base64_to_binary_details(data, options, last_chunk)
    cap = oracle.maximal_binary_length_from_base64(data)
    d = dev(data, misalign=misalign)
    _, v = out_buf(cap, torch.uint8, misalign=(misalign * 7) % 16)
    got = b.base64_to_binary_details(d, v, options, last_chunk)
    assert got[:2] == want[:2], (data[:80], len(data), options, last_chunk, got, want)
    check_guard(v, cap)
    if want[0] not in (7, 9):  # output_count is not pinned on INVALID_BASE64_CHARACTER / EXTRA_BITS
        assert got == want, (data[:80], options, last_chunk, got, want)
        assert v[:want[2]].cpu().numpy().tobytes() == wout.tobytes()
    if host_too:
        h = np.full(cap + GUARD, 0x5A, dtype=np.uint8)
        got = b.base64_to_binary_details(data, h, options, last_chunk)
        assert got[:2] == want[:2]
        assert (h[cap:] == 0x5A).all()
        if want[0] not in (7, 9):
            assert got == want and h[:want[2]].tobytes() == wout.tobytes()


# ------------------------------------------------------------------------------------------------------------
# golden vectors
# ------------------------------------------------------------------------------------------------------------
def test_golden_kat(b, oracle, golden):
    for v in golden["kat"]:
        data = bytes.fromhex(v["input"])
        f = v["func"]
        d = dev(data)
        if f == "validate_utf8":
            assert b.validate_utf8(d) == v["expect"], v
            assert b.validate_utf8(data) == v["expect"], v
        elif f == "validate_utf8_with_errors":
            assert list(b.validate_utf8_with_errors(d)) == v["expect"], v
            assert list(b.validate_utf8_with_errors(data)) == v["expect"], v
        elif f == "convert_utf8_to_utf16le_with_errors":
            n = oracle.utf16_length_from_utf8(data)
            _, o = out_buf(n, torch.int16)
            assert list(b.convert_utf8_to_utf16le_with_errors(d, o)) == v["expect"], v
            check_guard(o, n, 0x5A5A)
        elif f == "base64_to_binary":
            cap = oracle.maximal_binary_length_from_base64(data)
            _, o = out_buf(cap, torch.uint8)
            assert list(b.base64_to_binary(d, o, v["options"], v["last_chunk"])) == v["expect"], v
            if "output" in v:
                assert o[:v["expect"][1]].cpu().numpy().tobytes().hex() == v["output"], v


def test_golden_recorded(b, oracle, golden):
    for v in golden["recorded"]:
        data = bytes.fromhex(v["input"])
        if v["kind"] == "utf8":
            d = dev(data)
            assert list(b.validate_utf8_with_errors(d)) == v["validate"]
            assert b.count_utf8(d) == v["count_utf8"] and b.utf16_length_from_utf8(d) == v["utf16_length"]
            _, o = out_buf(v["utf16_length"], torch.int16)
            assert list(b.convert_utf8_to_utf16le_with_errors(d, o)) == v["to_utf16"]
            if v["to_utf16"][0] == 0:
                assert o[:v["utf16_length"]].cpu().numpy().tobytes().hex() == v["utf16_out"]
            _, o = out_buf(v["count_utf8"], torch.int32)
            assert list(b.convert_utf8_to_utf32_with_errors(d, o)) == v["to_utf32"]
            if v["to_utf32"][0] == 0:
                assert o[:v["count_utf8"]].cpu().numpy().tobytes().hex() == v["utf32_out"]
        elif v["kind"] == "utf16":
            d = dev(data).view(torch.int16)
            assert b.count_utf16le(d) == v["count_utf16le"] and b.utf8_length_from_utf16le(d) == v["utf8_length"]
            assert list(b.validate_utf16le_with_errors(d)) == v["validate"]
            _, o = out_buf(v["utf8_length"], torch.uint8)
            assert list(b.convert_utf16le_to_utf8_with_errors(d, o)) == v["to_utf8"]
            if v["to_utf8"][0] == 0:
                assert o[:v["utf8_length"]].cpu().numpy().tobytes().hex() == v["utf8_out"]
        else:
            d = dev(data)
            assert b.maximal_binary_length_from_base64(data) == v["maxlen"]
            for c in v["cases"]:
                _, o = out_buf(v["maxlen"], torch.uint8)
                got = b.base64_to_binary_details(d, o, c["options"], c["last_chunk"])
                assert list(got)[:2] == c["result"][:2], (data, c, got)
                if c["output"] is not None:
                    assert list(got) == c["result"], (data, c, got)
                    assert o[:got[2]].cpu().numpy().tobytes().hex() == c["output"]


# ------------------------------------------------------------------------------------------------------------
# seeded random / adversarial inputs, all sizes and misalignments
# ------------------------------------------------------------------------------------------------------------
SPECIAL = [0x20, 0x41, 0x7F, 0x80, 0x8F, 0x90, 0x9F, 0xA0, 0xBF, 0xC0, 0xC1, 0xC2, 0xDF, 0xE0, 0xE1, 0xEC, 0xED, 0xEE, 0xEF, 0xF0,
           0xF1, 0xF3, 0xF4, 0xF5, 0xF7, 0xF8, 0xFF]


def rand_text(rng, nchars):
    return "".join(chr(rng.choice([rng.randrange(0x20, 0x7f), rng.randrange(0xa0, 0x250), rng.randrange(0x4e00, 0xa000),
                                   rng.randrange(0x1f300, 0x1f650)])) for _ in range(nchars)).encode()


def test_utf8_small_random(b, oracle):
    rng = random.Random(101)
    for it in range(400):
        n = rng.choice([0, 1, 2, 3, 4, 15, 16, 17, 31, 32, 33, 63, 64, 65, 127, 128, 129, 255, 511, 512, 513, 2047, 2049, 5000])
        mode = it % 5
        if mode == 0:
            d = bytes(rng.choice(SPECIAL) for _ in range(n))
        elif mode == 1:
            d = bytes(rng.randrange(256) for _ in range(n))
        else:
            bb = bytearray(rand_text(rng, n // 2 + 1)[:n + 3])
            if mode == 3 and bb:
                bb[rng.randrange(len(bb))] = rng.choice(SPECIAL)
            if mode == 4 and bb:
                del bb[len(bb) - rng.randrange(min(4, len(bb))):]
            d = bytes(bb)
        run_utf8(b, oracle, d, misalign=rng.randrange(16), host_too=(it % 4 == 0))


def test_utf8_error_classes_at_tile_edges(b, oracle):
    """Each of the six error classes planted at granule / warp-chunk / tile boundaries of a
    multi-tile buffer of valid text; also the last byte (truncation)."""
    rng = random.Random(5)
    base = bytearray(rand_text(rng, 40000))  # ~100 KB: several 16 KiB tiles
    n = len(base)
    bad = {
        "HEADER_BITS": b"\xf8", "TOO_SHORT": b"\xe4\x20", "TOO_LONG": b"\x80", "OVERLONG": b"\xc0\x80",
        "TOO_LARGE": b"\xf4\x90\x80\x80", "SURROGATE": b"\xed\xa0\x80", "OVERLONG3": b"\xe0\x80\x80", "OVERLONG4": b"\xf0\x80\x80\x80",
        "TOO_LARGE2": b"\xf5\x80\x80\x80",
    }
    spots = [0, 13, 15, 16, 17, 511, 512, 513, 2047, 2048, 2049, 16383, 16384, 16385, 32767, 32768, 49151, 49152, n - 5, n - 1]
    for name, seq in bad.items():
        for pos in spots:
            d = bytearray(base)
            # move to a character start so the planted sequence is what the parser meets first
            p = min(pos, n - 1)
            while p > 0 and (d[p] & 0xC0) == 0x80:
                p -= 1
            d[p:p + len(seq)] = seq
            run_utf8(b, oracle, bytes(d), misalign=pos % 16, host_too=False)
    for cut in range(1, 4):  # buffer ends inside a 4-byte character
        d = bytes(base[:30000])
        while (d[-1] & 0xC0) == 0x80:
            d = d[:-1]
        d = d[:-1] + "\U0001F600".encode()[:4 - cut]
        run_utf8(b, oracle, d, host_too=True)


def test_utf8_medium_exact(b, oracle):
    from simdutf_b200 import synth
    for n, seed in ((1 << 20, 21), (3 * (1 << 20) + 12345, 22), (1 << 23, 23)):
        d = synth.mixed_utf8(n, seed=seed).numpy().tobytes()
        run_utf8(b, oracle, d, misalign=seed % 16, host_too=(n <= 1 << 21))
    a = synth.ascii_text((1 << 22) + 7, seed=1).numpy().tobytes()
    run_utf8(b, oracle, a, misalign=3, host_too=False)


def rand_units(rng, n, mode):
    u = []
    edge = [0x41, 0x7f, 0x80, 0x7ff, 0x800, 0xd7ff, 0xd800, 0xdbff, 0xdc00, 0xdfff, 0xe000, 0xffff]
    while len(u) < n:
        c = rng.randrange(6)
        if mode == 0: u.append(rng.choice(edge))
        elif c == 0: u.append(rng.randrange(0x80))
        elif c == 1: u.append(rng.randrange(0x80, 0x800))
        elif c == 2: u.append(rng.choice([rng.randrange(0x800, 0xd800), rng.randrange(0xe000, 0x10000)]))
        elif c == 3: u += [rng.randrange(0xd800, 0xdc00), rng.randrange(0xdc00, 0xe000)]
        elif mode == 2 and rng.random() < 0.05: u.append(rng.randrange(0xd800, 0xe000))
        else: u.append(0x20)
    return np.array(u, dtype=np.uint16)


def test_utf16_random_and_surrogate_errors(b, oracle):
    rng = random.Random(202)
    for it in range(200):
        n = rng.choice([0, 1, 2, 7, 8, 9, 15, 16, 17, 255, 256, 257, 1000, 4095, 4097, 9000])
        run_utf16(b, oracle, rand_units(rng, n, it % 3), misalign=rng.randrange(8), host_too=(it % 4 == 0))
    from simdutf_b200 import synth
    base = synth.mixed_utf16le(100000, seed=31).numpy().view(np.uint16)
    run_utf16(b, oracle, base, host_too=True)
    # lone low / lone high / low-low / pair-then-high at tile edges (reference tests/convert_utf16le_to_utf8_with_errors_tests.cpp:110-212)
    for pos in (0, 7, 8, 255, 256, 8191, 8192, 8193, 16383, 16384, len(base) - 1):
        for planted in ([0xDC00], [0xD800], [0xDC00, 0xDC00], [0xD800, 0xD800], [0xD800, 0x41]):
            u = base.copy()
            p = pos
            while p > 0 and (u[p] & 0xFC00) == 0xDC00:
                p -= 1
            u[p:p + len(planted)] = planted[:len(u) - p]
            run_utf16(b, oracle, u, misalign=pos % 8, host_too=False)


ABC = b"ABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz0123456789+/-_"


def test_base64_random_all_options(b, oracle):
    rng = random.Random(303)
    fixed = [b"", b" ", b"=", b"==", b"A", b"AA", b"AAA", b"AAAA", b"AA=", b"AA==", b"AAA=", b"AAA==", b"AAAA=", b"A=", b"QQ===",
             b"AA=A", b"AAAA*AAA", b"AAAA AAAA\r\n", b"AAAAA", b" A A = = ", b"AAA\x80", b"AA-A", b"AA+A"]
    for it in range(260):
        if it < len(fixed):
            d = fixed[it]
        else:
            n = rng.choice([3, 4, 5, 15, 16, 17, 63, 64, 65, 66, 100, 1000, 4095, 4096, 4097, 20000, 40000])
            out = bytearray()
            for _ in range(n):
                x = rng.random()
                if x < 0.85: out.append(rng.choice(ABC[:64] if it % 2 else ABC[:62] + ABC[64:]))
                elif x < 0.97: out.append(rng.choice(b" \t\n\r\x0c"))
                elif it % 5 == 0: out.append(rng.choice(b"=*\x80\xff.\x0b"))
                else: out.append(rng.choice(ABC[:62]))
            d = bytes(out) + rng.choice([b"", b"=", b"==", b" = = ", b"= ", b"=\n=", b" ", b"   \r\n" * 50])
        for opt in ((0, 1, 2, 3, 4, 5, 8, 12) if it % 3 == 0 or it < len(fixed) else (it % 2,)):
            for lc in (0, 1, 2):
                run_b64(b, oracle, d, opt, lc, misalign=rng.randrange(16), host_too=(it % 8 == 0))


def test_base64_large_with_whitespace(b, oracle):
    from simdutf_b200 import synth
    for url in (False, True):
        text, payload = synth.base64_text(3 << 20, seed=41 + url, url=url)
        data = text.numpy().tobytes()
        run_b64(b, oracle, data, options=1 if url else 0, misalign=5, host_too=True)
        cap = oracle.maximal_binary_length_from_base64(data)
        _, o = out_buf(cap, torch.uint8)
        e, i, n = b.base64_to_binary_details(dev(data), o, 1 if url else 0, 0)
        assert e == 0 and n == payload.numel() and o[:n].cpu().numpy().tobytes() == payload.numpy().tobytes()
        # one bad character deep inside; a dangling single sextet at the end
        for q in (0, 77, 16384, 16385, 1 << 20, len(data) - 9):
            bad = bytearray(data); bad[q] = ord("*")
            run_b64(b, oracle, bytes(bad), options=1 if url else 0, host_too=False)
        stripped = data.rstrip(b"=\r\n \t")
        for k in range(1, 5):  # a dangling single sextet at the very end -> BASE64_INPUT_REMAINDER
            cand = stripped + b"A" * k
            if oracle.base64_to_binary_details(cand, 1 if url else 0, 0)[0][0] == 8:
                break
        assert oracle.base64_to_binary_details(cand, 1 if url else 0, 0)[0][0] == 8
        run_b64(b, oracle, cand, options=1 if url else 0, host_too=False)


def test_base64_single_pass_structure(b, oracle):
    """The single-pass decoder's own cases: warp-tiles without any whitespace (the 128-bit staging path), next to tiles
    with whitespace, tiles of nothing but whitespace (their successors fetch the carried sextets from far back), every
    phase of the tile's first quantum, every length of the trailing partial quantum, sizes around the CTA-tile."""
    import base64 as pyb64
    rng = random.Random(909)
    raw = bytes(rng.getrandbits(8) for _ in range(300_000))
    dense = pyb64.b64encode(raw)  # 400 000 characters, no whitespace, no padding
    cta_tile = 7 * 2048
    for n in (64, 2047, 2048, 2049, 4096, cta_tile - 1, cta_tile, cta_tile + 1, 3 * cta_tile + 5, 148 * 4 * cta_tile // 8 + 3, 399_999, 400_000):
        for mis in (0, 1, 31):
            run_b64(b, oracle, dense[:n], 0, 0, misalign=mis, host_too=False)
    for lc in (0, 1, 2):  # trailing partial quanta of 1, 2, 3 sextets behind dense tiles, with and without padding
        for tail in (b"", b"Q", b"QQ", b"QQ=", b"QQ==", b"QQQ", b"QQQ=", b"QUE", b"QUE=", b"QR", b"QR=="):
            run_b64(b, oracle, dense[:40_000] + tail, 0, lc, misalign=3, host_too=False)
    # dense and sparse tiles interleaved; a run of whitespace longer than several tiles; whitespace that shifts the phase
    for k in range(1, 6):
        mixed = dense[:10_000 + k] + b" " * k + dense[10_000 + k:60_000] + b"\r\n" * 5000 + dense[60_000:200_000 - k] + b"\t" + dense[200_000 - k:300_000]
        run_b64(b, oracle, mixed, 0, 0, misalign=k, host_too=(k == 1))
        run_b64(b, oracle, mixed, 4, 0, misalign=k, host_too=False)
    # an invalid character in a dense tile, in a sparse tile, in the first and in the last tile
    for q in (0, 5000, 10_002, 75_000, 299_999):
        bad = bytearray(dense[:10_000] + b" " + dense[10_000:300_000]); bad[q] = ord("*")
        run_b64(b, oracle, bytes(bad), 0, 0, host_too=False)
        run_b64(b, oracle, bytes(bad), 4, 0, host_too=False)  # tolerant: the character is skipped, phases shift


def test_bitplane_edge_paths(b, oracle):
    """Paths the bit-plane kernels take only on unusual inputs: all-ASCII tiles next to non-ASCII ones, lanes that
    emit fewer than one vector (unit-by-unit copy-out), a buffer that starts with a continuation byte, UTF-16
    buffers that end with a high surrogate exactly at a tile edge, base64 tiles that are all whitespace (the
    backward carry scan crosses them) or hold 1-3 sextets."""
    rng = random.Random(77)
    # UTF-8: ASCII runs of tile size glued to mixed text, at every alignment class
    for it in range(12):
        parts = []
        for _ in range(6):
            parts.append(bytes(rng.randrange(0x20, 0x7f) for _ in range(rng.choice([1, 63, 64, 2047, 2048, 2049, 4100]))))
            parts.append(rand_text(rng, rng.choice([1, 5, 700, 3000])))
        run_utf8(b, oracle, b"".join(parts), misalign=rng.randrange(16), host_too=(it == 0))
    # invalid from the first byte / nothing but continuation bytes / long runs of them inside valid text
    run_utf8(b, oracle, b"\x80" + rand_text(rng, 5000), misalign=1, host_too=False)
    run_utf8(b, oracle, b"\x80" * 10000, misalign=7, host_too=False)
    run_utf8(b, oracle, b"\xbf" * 3 + b"abc", host_too=False)
    t = bytearray(rand_text(rng, 9000))
    t[5000:9000] = b"\x80" * 4000
    run_utf8(b, oracle, bytes(t), misalign=9, host_too=False)
    run_utf8(b, oracle, rand_text(rng, 3000) + b"\xf0\x9f\x98", misalign=2, host_too=False)
    # 4-byte characters only (two units per character), 3-byte only, 2-byte only: regular strides
    for ch in ("\U0001F600", "\u4e2d", "\u00e9", "a"):
        run_utf8(b, oracle, (ch * 9001).encode(), misalign=rng.randrange(16), host_too=False)
    # UTF-16: ASCII tiles, pairs only, a high surrogate as the very last unit at / around a 1024-unit tile edge
    for n in (1023, 1024, 1025, 2048, 4096):
        for mis in (0, 3):
            u = np.array([rng.randrange(0x20, 0x7f) for _ in range(n)], dtype=np.uint16)
            run_utf16(b, oracle, u, misalign=mis, host_too=False)
            u2 = u.copy(); u2[-1] = 0xD83D
            run_utf16(b, oracle, u2, misalign=mis, host_too=False)
            u3 = u.copy(); u3[0] = 0xDC00
            run_utf16(b, oracle, u3, misalign=mis, host_too=False)
    pairs = np.array([0xD83D, 0xDE00] * 3000, dtype=np.uint16)
    run_utf16(b, oracle, pairs, misalign=1, host_too=False)
    run_utf16(b, oracle, np.array([0x4E2D] * 5000, dtype=np.uint16), misalign=5, host_too=False)
    # base64: whitespace deserts between sextets, sextets sprinkled one per tile, whitespace only
    for opt in (0, 1, 4):
        for lc in (0, 1, 2):
            d = b"QUJD" + b" " * 9000 + b"R" + b"\n" * 5000 + b"EVG" + b"\r\n" * 3000 + b"R0g="
            run_b64(b, oracle, d, opt, lc, misalign=rng.randrange(16), host_too=False)
            run_b64(b, oracle, b" \t\r\n" * 3000, opt, lc, host_too=False)
            sparse = bytearray(b" " * 20000)
            for k, q in enumerate(range(100, 20000, 2100)):
                sparse[q] = ABC[k % 62]
            run_b64(b, oracle, bytes(sparse), opt, lc, misalign=3, host_too=False)
            dense = bytes(rng.choice(ABC[:62]) for _ in range(8190)) + rng.choice([b"", b"A", b"AA", b"AA=", b"A=="])
            run_b64(b, oracle, dense, opt, lc, misalign=rng.randrange(16), host_too=False)


def test_single_pass_transcoder_structure(b, oracle):
    """K3 (k_utf8_transcode_v3): sizes around its CTA-tile (16 warp-tiles of 3 KiB = 48 KiB; UTF-32: 12 x 2 KiB), around
    one wave of the persistent grid (one CTA per SM: 148) and a few waves beyond it — ticket hand-out past the first wave, the four-slot
    hand-off rings wrapping, look-backs longer than one 128-descriptor window — with whole-tile ASCII runs (copied out
    without staging) between mixed text, every output alignment class (the deferred copy-out realigns by words), an
    error in the last wave, and a buffer that is nothing but sentinel tickets for most CTAs."""
    rng = random.Random(4711)
    tile, grid = 16 * 3072, 148
    base = rand_text(rng, 1024 * 1024)
    for n in (1, 2047, 2048, 2049, tile - 1, tile, tile + 1, 2 * tile + 17, 12 * 2048 + 3, grid * tile - 5, grid * tile + 3072, 3 * grid * tile + 12345):
        data = bytearray((base * (n // len(base) + 1))[:n])
        while data and (data[-1] & 0xC0) == 0x80:   # do not end inside a character ...
            data.pop()
        if data and data[-1] >= 0xC0:               # ... nor on a lead
            data.pop()
        # whole-tile ASCII runs at tile-aligned and unaligned places
        for k in range(3 * tile + 5, len(data) - 4 * tile, 37 * tile + 1000):
            data[k:k + 3 * tile] = bytes(0x41 + (i % 26) for i in range(3 * tile))
        data = bytes(data)
        for mis in (0, 1, 5, 15):
            run_utf8(b, oracle, _repair(data), misalign=mis, host_too=False)
    big = bytearray(_repair(bytes((base * 9)[: 3 * grid * tile + 999])))
    big[len(big) - 3 * tile - 7] = 0xFF            # an error in the last wave
    run_utf8(b, oracle, bytes(big), misalign=3, host_too=False)


def _repair(data: bytes) -> bytes:
    """Make a spliced byte string valid UTF-8 again: decode leniently and re-encode."""
    return data.decode("utf-8", errors="ignore").encode("utf-8")


def test_utf16be_twins(b, oracle):
    """the UTF-16BE twins and change_endianness_utf16 against the oracle — random text, every
    misalignment class, surrogate errors at tile edges, the host path."""
    rng = random.Random(2024)
    for it in range(60):
        n = rng.choice([0, 1, 2, 7, 8, 9, 31, 32, 33, 1023, 1024, 1025, 5000, 70000])
        d = rand_text(rng, n // 2 + 1) if it % 3 else bytes(rng.choice(SPECIAL) for _ in range(n))
        mis = rng.randrange(16)
        want, wout = oracle.convert_utf8_to_utf16be_with_errors(d)
        n16 = oracle.utf16_length_from_utf8(d)
        dd = dev(d, misalign=mis)
        _, v16 = out_buf(n16, torch.int16, misalign=mis % 8)
        assert b.convert_utf8_to_utf16be_with_errors(dd, v16) == want, (d[:64].hex(), len(d), mis)
        if want[0] == 0:
            assert v16[:want[1]].cpu().numpy().view(np.uint16).tobytes() == wout.tobytes()
            check_guard(v16, want[1], 0x5A5A)
            if it % 5 == 0:
                h = np.zeros(n16 + 4, dtype=np.uint16)
                assert b.convert_utf8_to_utf16be_with_errors(d, h) == want and h[:want[1]].tobytes() == wout.tobytes()
    for it in range(60):
        n = rng.choice([0, 1, 2, 7, 8, 9, 1023, 1024, 1025, 4096, 9000, 40000])
        u = rand_units(rng, n, it % 3)
        if it % 4 == 0 and n:
            u[-1] = 0xD800 + rng.randrange(0x400)  # ends with a high surrogate
        be = u.byteswap()
        mis = rng.randrange(8)
        want, wout = oracle.convert_utf16be_to_utf8_with_errors(be)
        dd = dev(be.view(np.uint8), dtype=torch.uint8, misalign=2 * mis).view(torch.int16)
        assert b.count_utf16be(dd) == oracle.count_utf16be(be)
        assert b.utf8_length_from_utf16be(dd) == oracle.utf8_length_from_utf16be(be)
        assert b.validate_utf16be_with_errors(dd) == oracle.validate_utf16be_with_errors(be)
        n8 = oracle.utf8_length_from_utf16be(be)
        _, v8 = out_buf(n8, torch.uint8, misalign=mis)
        assert b.convert_utf16be_to_utf8_with_errors(dd, v8) == want, (u[:16], len(u), mis)
        if want[0] == 0:
            assert v8[:want[1]].cpu().numpy().tobytes() == wout.tobytes()
            check_guard(v8, want[1], 0x5A)
        for omis in (0, mis, (mis + 3) % 8):
            _, sw = out_buf(len(be), torch.int16, misalign=omis)
            b.change_endianness_utf16(dd, sw)
            assert sw[:len(be)].cpu().numpy().view(np.uint16).tobytes() == u.tobytes()
            check_guard(sw, len(be), 0x5A5A)
        if it % 6 == 0:
            assert b.count_utf16be(be) == oracle.count_utf16be(be)
            h8 = np.zeros(n8 + 4, dtype=np.uint8)
            assert b.convert_utf16be_to_utf8_with_errors(be, h8) == want
            hs = np.zeros(len(be) + 1, dtype=np.uint16)
            b.change_endianness_utf16(be, hs)
            assert hs[:len(be)].tobytes() == u.tobytes()
    # a lone surrogate planted at tile edges of a multi-tile buffer
    base = rand_units(rng, 6000, 1)
    for pos in (0, 1023, 1024, 1025, 2047, 2048, len(base) - 1):
        for planted in (0xDC00, 0xD800):
            u = base.copy()
            u[pos] = planted
            be = u.byteswap()
            dd = dev(be.view(np.uint8), dtype=torch.uint8, misalign=2 * (pos % 8)).view(torch.int16)
            assert b.validate_utf16be_with_errors(dd) == oracle.validate_utf16be_with_errors(be)
            _, v8 = out_buf(oracle.utf8_length_from_utf16be(be), torch.uint8)
            assert b.convert_utf16be_to_utf8_with_errors(dd, v8) == oracle.convert_utf16be_to_utf8_with_errors(be)[0]
    # 256 MiB round trip: UTF-8 -> UTF-16BE -> swap == UTF-16LE, UTF-16BE -> UTF-8 == input
    from simdutf_b200 import synth
    d = synth.mixed_utf8(1 << 28, seed=12, device="cuda")
    units = b.utf16_length_from_utf8(d)
    ube = torch.empty(units, dtype=torch.int16, device="cuda")
    ule = torch.empty(units, dtype=torch.int16, device="cuda")
    assert b.convert_utf8_to_utf16be_with_errors(d, ube) == (0, units)
    assert b.convert_utf8_to_utf16le_with_errors(d, ule) == (0, units)
    sw = torch.empty_like(ube)
    b.change_endianness_utf16(ube, sw)
    assert torch.equal(sw, ule)
    assert b.count_utf16be(ube) == b.count_utf16le(ule) and b.utf8_length_from_utf16be(ube) == d.numel()
    back = torch.empty_like(d)
    assert b.convert_utf16be_to_utf8_with_errors(ube, back) == (0, d.numel())
    assert torch.equal(back, d)


def rand_cps(rng, n, mode):
    """UTF-32 code points: mode 0 ASCII-heavy, 1 mixed widths, 2 mixed + occasional surrogates / out-of-range."""
    r = np.random.default_rng(rng.randrange(1 << 30))
    if mode == 0:
        a = r.integers(0x20, 0x7F, size=n, dtype=np.uint32)
        k = r.random(n) < 0.02
        a[k] = r.integers(0x80, 0x800, size=int(k.sum()), dtype=np.uint32)
        return a
    cls = r.integers(0, 4, size=n)
    a = np.where(cls == 0, r.integers(0, 0x80, size=n),
        np.where(cls == 1, r.integers(0x80, 0x800, size=n),
        np.where(cls == 2, r.integers(0xE000, 0x10000, size=n), r.integers(0x10000, 0x110000, size=n)))).astype(np.uint32)
    edge = r.random(n) < 0.05
    a[edge] = r.choice(np.array([0x7F, 0x80, 0x7FF, 0x800, 0xD7FF, 0xE000, 0xFFFF, 0x10000, 0x10FFFF], dtype=np.uint32),
                       size=int(edge.sum()))
    if mode == 2 and n:
        for _ in range(rng.randrange(1, 3)):
            a[rng.randrange(n)] = rng.choice([0xD800, 0xDBFF, 0xDC00, 0xDFFF, 0x110000, 0xFFFFFFFF, 0x80000000])
    return a


def test_utf32_family(b, oracle):
    """SURVEY.md §8f rank 1 (second part): validate_utf32_with_errors, utf8/utf16_length_from_utf32,
    convert_utf32_to_utf8 / _utf16le / _utf16be, convert_utf16le/be_to_utf32 against the oracle — every width
    mix, surrogate / too-large errors (first error wins), misaligned inputs and outputs, the host path, and a
    256 MiB round trip."""
    rng = random.Random(3232)
    sizes = [0, 1, 2, 3, 7, 8, 9, 15, 16, 17, 31, 32, 33, 255, 256, 257, 511, 512, 513, 1023, 1024, 1025, 3000, 20000, 70001]
    for it in range(90):
        n = rng.choice(sizes)
        a = rand_cps(rng, n, it % 3)
        mis = rng.randrange(8)
        dd = dev(a.view(np.uint8), misalign=4 * mis).view(torch.int32)
        assert b.validate_utf32_with_errors(dd) == oracle.validate_utf32_with_errors(a), (a[:16], n, mis)
        assert b.utf8_length_from_utf32(dd) == oracle.utf8_length_from_utf32(a)
        assert b.utf16_length_from_utf32(dd) == oracle.utf16_length_from_utf32(a)
        n8, n16 = oracle.utf8_length_from_utf32(a), oracle.utf16_length_from_utf32(a)
        want, wout = oracle.convert_utf32_to_utf8_with_errors(a)
        _, v8 = out_buf(n8, torch.uint8, misalign=rng.randrange(16))
        assert b.convert_utf32_to_utf8_with_errors(dd, v8) == want, (a[:16], n, mis)
        if want[0] == 0:
            assert v8[:want[1]].cpu().numpy().tobytes() == wout.tobytes(), (n, mis, it)
            check_guard(v8, want[1])
        for be in (False, True):
            want, wout = oracle.convert_utf32_to_utf16_with_errors(a, be)
            _, v16 = out_buf(n16, torch.int16, misalign=rng.randrange(8))
            fn = b.convert_utf32_to_utf16be_with_errors if be else b.convert_utf32_to_utf16le_with_errors
            assert fn(dd, v16) == want, (a[:16], n, mis, be)
            if want[0] == 0:
                assert v16[:want[1]].cpu().numpy().view(np.uint16).tobytes() == wout.tobytes(), (n, mis, be, it)
                check_guard(v16, want[1])
        if it % 6 == 0:  # host path
            assert b.validate_utf32_with_errors(a) == oracle.validate_utf32_with_errors(a)
            assert b.utf8_length_from_utf32(a) == n8 and b.utf16_length_from_utf32(a) == n16
            want, wout = oracle.convert_utf32_to_utf8_with_errors(a)
            h8 = np.zeros(n8 + 4, dtype=np.uint8)
            assert b.convert_utf32_to_utf8_with_errors(a, h8) == want
            if want[0] == 0:
                assert h8[:want[1]].tobytes() == wout.tobytes()
            want, wout = oracle.convert_utf32_to_utf16_with_errors(a, True)
            h16 = np.zeros(n16 + 4, dtype=np.uint16)
            assert b.convert_utf32_to_utf16be_with_errors(a, h16) == want
            if want[0] == 0:
                assert h16[:want[1]].tobytes() == wout.tobytes()
    # UTF-16 -> UTF-32, LE and BE, valid and with surrogate errors (pair split across a tile edge included)
    for it in range(80):
        n = rng.choice(sizes)
        u = rand_units(rng, n, it % 3)
        if it % 4 == 0 and n:
            u[-1] = 0xD800 + rng.randrange(0x400)
        if it % 5 == 0 and n > 1025:
            u[1023], u[1024] = 0xD800 + rng.randrange(0x400), 0xDC00 + rng.randrange(0x400)
        for be in (False, True):
            src = u.byteswap() if be else u
            mis = rng.randrange(8)
            want, wout = oracle.convert_utf16_to_utf32_with_errors(src, be)
            dd = dev(src.view(np.uint8), misalign=2 * mis).view(torch.int16)
            n32 = oracle.count_utf16be(src) if be else oracle.count_utf16le(src)
            _, v32 = out_buf(n32, torch.int32, misalign=rng.randrange(4))
            fn = b.convert_utf16be_to_utf32_with_errors if be else b.convert_utf16le_to_utf32_with_errors
            assert fn(dd, v32) == want, (u[:16], n, mis, be)
            if want[0] == 0:
                assert v32[:want[1]].cpu().numpy().view(np.uint32).tobytes() == wout.tobytes(), (n, mis, be, it)
                check_guard(v32, want[1])
            if it % 8 == 0:
                h32 = np.zeros(n32 + 4, dtype=np.uint32)
                assert fn(src, h32) == want
                if want[0] == 0:
                    assert h32[:want[1]].tobytes() == wout.tobytes()
    # an invalid code point planted at tile edges: first error wins, earlier errors beat later ones
    base = rand_cps(rng, 5000, 1)
    for pos in (0, 511, 512, 513, 1023, 1024, 4999):
        for planted in (0xD800, 0xDFFF, 0x110000, 0xFFFFFFFF):
            a = base.copy()
            a[pos] = planted
            if pos < 4000:
                a[4500] = 0xDC00  # a later error must not win
            dd = dev(a.view(np.uint8)).view(torch.int32)
            assert b.validate_utf32_with_errors(dd) == oracle.validate_utf32_with_errors(a)
            _, v8 = out_buf(oracle.utf8_length_from_utf32(a), torch.uint8)
            assert b.convert_utf32_to_utf8_with_errors(dd, v8) == oracle.convert_utf32_to_utf8_with_errors(a)[0]
            _, v16 = out_buf(oracle.utf16_length_from_utf32(a), torch.int16)
            assert b.convert_utf32_to_utf16le_with_errors(dd, v16) == oracle.convert_utf32_to_utf16_with_errors(a)[0]
    # 256 MiB of UTF-8 -> UTF-32 (the UTF-8 kernel) -> UTF-8 / UTF-16 (these kernels) -> compare with the direct paths
    from simdutf_b200 import synth
    d = synth.mixed_utf8(1 << 28, seed=32, device="cuda")
    cps = b.count_utf8(d)
    u32 = torch.empty(cps, dtype=torch.int32, device="cuda")
    assert b.convert_utf8_to_utf32_with_errors(d, u32) == (0, cps)
    assert b.validate_utf32_with_errors(u32) == (0, cps)
    assert b.utf8_length_from_utf32(u32) == d.numel()
    units = b.utf16_length_from_utf8(d)
    assert b.utf16_length_from_utf32(u32) == units
    back8 = torch.empty(d.numel(), dtype=torch.uint8, device="cuda")
    assert b.convert_utf32_to_utf8_with_errors(u32, back8) == (0, d.numel())
    assert torch.equal(back8, d)
    del back8
    ule = torch.empty(units, dtype=torch.int16, device="cuda")
    assert b.convert_utf8_to_utf16le_with_errors(d, ule) == (0, units)
    got = torch.empty(units, dtype=torch.int16, device="cuda")
    assert b.convert_utf32_to_utf16le_with_errors(u32, got) == (0, units)
    assert torch.equal(got, ule)
    assert b.convert_utf32_to_utf16be_with_errors(u32, got) == (0, units)
    sw = torch.empty_like(got)
    b.change_endianness_utf16(got, sw)
    assert torch.equal(sw, ule)
    back32 = torch.empty(cps, dtype=torch.int32, device="cuda")
    assert b.convert_utf16be_to_utf32_with_errors(got, back32) == (0, cps)
    assert torch.equal(back32, u32)
    assert b.convert_utf16le_to_utf32_with_errors(ule, back32) == (0, cps)
    assert torch.equal(back32, u32)


def test_latin1_family(b, oracle):
    """validate_ascii_with_errors, utf8_length_from_latin1, latin1_length_from_utf8,
    convert_latin1_to_utf8 / _utf16le / _utf16be / _utf32, convert_utf8 / utf16le / utf16be / utf32_to_latin1
    against the oracle — all byte values, every error class of the UTF-8 walk, misaligned inputs and outputs,
    the host path, and a 256 MiB round trip."""
    rng = random.Random(8859)
    sizes = [0, 1, 2, 3, 4, 15, 16, 17, 31, 32, 33, 63, 64, 65, 127, 128, 129, 2047, 2048, 2049, 4096, 10000, 70001, 300000]
    u8pool = [0x00, 0x41, 0x7f, 0x80, 0xbf, 0xc0, 0xc1, 0xc2, 0xc3, 0xc4, 0xdf, 0xe0, 0xef, 0xf0, 0xf7, 0xf8, 0xff]
    for it in range(90):
        n = rng.choice(sizes)
        r = np.random.default_rng(rng.randrange(1 << 30))
        a = r.integers(0, 0x80, size=n, dtype=np.uint8)
        if it % 3:
            k = r.random(n) < (0.02 if it % 3 == 1 else 0.4)
            a[k] = r.integers(0x80, 0x100, size=int(k.sum()), dtype=np.uint8)
        d = a.tobytes()
        mis = rng.randrange(16)
        dd = dev(d, misalign=mis)
        assert b.validate_ascii_with_errors(dd) == oracle.validate_ascii_with_errors(d), (n, mis, it)
        assert b.utf8_length_from_latin1(dd) == oracle.utf8_length_from_latin1(d)
        want = oracle.convert_latin1_to_utf8(d)
        _, v8 = out_buf(len(want), torch.uint8, misalign=rng.randrange(16))
        assert b.convert_latin1_to_utf8(dd, v8) == (0, len(want)), (n, mis, it)
        assert v8[:len(want)].cpu().numpy().tobytes() == want.tobytes(), (n, mis, it)
        check_guard(v8, len(want))
        for be in (False, True):
            w16 = oracle.convert_latin1_to_utf16(d, be)
            _, v16 = out_buf(n, torch.int16, misalign=rng.randrange(8))
            fn = b.convert_latin1_to_utf16be if be else b.convert_latin1_to_utf16le
            assert fn(dd, v16) == (0, n)
            assert v16[:n].cpu().numpy().view(np.uint16).tobytes() == w16.tobytes(), (n, mis, be, it)
            check_guard(v16, n)
        w32 = oracle.convert_latin1_to_utf32(d)
        _, v32 = out_buf(n, torch.int32, misalign=rng.randrange(4))
        assert b.convert_latin1_to_utf32(dd, v32) == (0, n)
        assert v32[:n].cpu().numpy().view(np.uint32).tobytes() == w32.tobytes(), (n, mis, it)
        check_guard(v32, n)
        # and back: the UTF-8 we just made is valid Latin-1-range UTF-8
        u8 = want.tobytes()
        d8 = dev(u8, misalign=rng.randrange(16))
        assert b.latin1_length_from_utf8(d8) == n
        _, vl = out_buf(n, torch.uint8, misalign=rng.randrange(16))
        assert b.convert_utf8_to_latin1_with_errors(d8, vl) == (0, n), (n, it)
        assert vl[:n].cpu().numpy().tobytes() == d
        check_guard(vl, n)
        if it % 6 == 0:  # host path
            assert b.validate_ascii_with_errors(d) == oracle.validate_ascii_with_errors(d)
            assert b.utf8_length_from_latin1(d) == len(want)
            h8 = np.zeros(len(want) + 4, dtype=np.uint8)
            assert b.convert_latin1_to_utf8(d, h8) == (0, len(want)) and h8[:len(want)].tobytes() == want.tobytes()
            h16 = np.zeros(n + 4, dtype=np.uint16)
            assert b.convert_latin1_to_utf16be(d, h16) == (0, n) and h16[:n].tobytes() == oracle.convert_latin1_to_utf16(d, True).tobytes()
            hl = np.zeros(n + 4, dtype=np.uint8)
            assert b.convert_utf8_to_latin1_with_errors(u8, hl) == (0, n) and hl[:n].tobytes() == d
    # UTF-8 -> Latin-1 error classes: damaged Latin-1-range UTF-8 and byte soup from the class edges
    for it in range(160):
        n = rng.choice(sizes[:20] if it < 120 else [6000, 10000, 70001])  # the larger ones have interior tiles (SWAR screen)
        if it % 2:
            t = bytearray("".join(chr(rng.randrange(0x100) if rng.random() < 0.4 else rng.randrange(0x80)) for _ in range(n)).encode())
            for _ in range(rng.randrange(1, 3)):
                if t:
                    t[rng.randrange(len(t))] = rng.choice(u8pool)
            if it >= 120 and it % 4 == 1:  # exactly one damaged byte, deep inside
                t = bytearray("".join(chr(rng.randrange(0x100) if rng.random() < 0.4 else rng.randrange(0x80)) for _ in range(n)).encode())
                t[rng.randrange(2100, len(t) - 2100)] = rng.choice(u8pool[3:])
        else:
            t = bytearray(rng.choice(u8pool) if rng.random() < 0.05 else 0x41 for _ in range(n))
            if n > 2049 and it % 4 == 0:
                t[2047:2049] = b"\xc3\xa9"  # a 2-byte character across the tile edge
        t = bytes(t)
        mis = rng.randrange(16)
        want, wout = oracle.convert_utf8_to_latin1_with_errors(t)
        dd = dev(t, misalign=mis)
        _, vl = out_buf(oracle.count_utf8(t), torch.uint8, misalign=rng.randrange(16))
        assert b.convert_utf8_to_latin1_with_errors(dd, vl) == want, (t[:64].hex(), len(t), mis, it)
        if want[0] == 0:
            assert vl[:want[1]].cpu().numpy().tobytes() == wout.tobytes()
            check_guard(vl, want[1])
        if it % 10 == 0:
            hl = np.zeros(len(t) + 4, dtype=np.uint8)
            assert b.convert_utf8_to_latin1_with_errors(t, hl) == want
    # UTF-16 / UTF-32 -> Latin-1 with a too-large element planted (first one wins)
    for it in range(60):
        n = rng.choice(sizes)
        r = np.random.default_rng(rng.randrange(1 << 30))
        a16 = r.integers(0, 0x100, size=n, dtype=np.uint16)
        a32 = a16.astype(np.uint32)
        if it % 2 and n:
            for _ in range(rng.randrange(1, 3)):
                k = rng.randrange(n)
                a16[k] = rng.choice([0x100, 0x7ff, 0xd800, 0xff00, 0xffff])
                a32[k] = rng.choice([0x100, 0xffff, 0x10ffff, 0xffffff41, 0x80000000])
        for be in (False, True):
            src = a16.byteswap() if be else a16
            mis = rng.randrange(8)
            want, wout = oracle.convert_utf16_to_latin1_with_errors(src, be)
            dd = dev(src.view(np.uint8), misalign=2 * mis).view(torch.int16)
            _, vl = out_buf(n, torch.uint8, misalign=rng.randrange(16))
            fn = b.convert_utf16be_to_latin1_with_errors if be else b.convert_utf16le_to_latin1_with_errors
            assert fn(dd, vl) == want, (n, mis, be, it)
            if want[0] == 0:
                assert vl[:n].cpu().numpy().tobytes() == wout.tobytes()
                check_guard(vl, n)
            if it % 10 == 0:
                hl = np.zeros(n + 4, dtype=np.uint8)
                assert fn(src, hl) == want
        want, wout = oracle.convert_utf32_to_latin1_with_errors(a32)
        dd = dev(a32.view(np.uint8), misalign=4 * rng.randrange(4)).view(torch.int32)
        _, vl = out_buf(n, torch.uint8, misalign=rng.randrange(16))
        assert b.convert_utf32_to_latin1_with_errors(dd, vl) == want, (n, it)
        if want[0] == 0:
            assert vl[:n].cpu().numpy().tobytes() == wout.tobytes()
            check_guard(vl, n)
    # 256 MiB of Latin-1 (30 % high bytes): -> UTF-8 -> validate as UTF-8 -> back; -> UTF-16LE -> back; -> UTF-32 -> back
    n = 1 << 28
    g = torch.Generator(device="cuda").manual_seed(59)
    lat = torch.randint(0, 0x80, (n,), dtype=torch.uint8, device="cuda", generator=g)
    hi = torch.rand(n, device="cuda", generator=g) < 0.3
    lat = torch.where(hi, lat | 0x80, lat)
    del hi
    n8 = b.utf8_length_from_latin1(lat)
    assert n8 == n + int((lat >= 0x80).sum())
    assert b.validate_ascii_with_errors(lat) == (5, int(torch.nonzero(lat >= 0x80)[0]))
    u8 = torch.empty(n8, dtype=torch.uint8, device="cuda")
    assert b.convert_latin1_to_utf8(lat, u8) == (0, n8)
    assert b.validate_utf8_with_errors(u8) == (0, n8)
    assert b.latin1_length_from_utf8(u8) == n
    back = torch.empty(n, dtype=torch.uint8, device="cuda")
    assert b.convert_utf8_to_latin1_with_errors(u8, back) == (0, n)
    assert torch.equal(back, lat)
    del u8
    u16 = torch.empty(n, dtype=torch.int16, device="cuda")
    assert b.convert_latin1_to_utf16le(lat, u16) == (0, n)
    assert torch.equal(u16, lat.to(torch.int16))
    back.zero_()
    assert b.convert_utf16le_to_latin1_with_errors(u16, back) == (0, n)
    assert torch.equal(back, lat)
    assert b.convert_latin1_to_utf16be(lat, u16) == (0, n)
    back.zero_()
    assert b.convert_utf16be_to_latin1_with_errors(u16, back) == (0, n)
    assert torch.equal(back, lat)
    del u16
    u32 = torch.empty(n, dtype=torch.int32, device="cuda")
    assert b.convert_latin1_to_utf32(lat, u32) == (0, n)
    back.zero_()
    assert b.convert_utf32_to_latin1_with_errors(u32, back) == (0, n)
    assert torch.equal(back, lat)


def test_well_formed_and_detect(b, oracle):
    """to_well_formed_utf16le/be (also in place) and detect_encodings against the oracle."""
    rng = random.Random(44)
    sizes = [0, 1, 2, 7, 8, 9, 15, 16, 17, 1023, 1024, 1025, 5000, 70001]
    for it in range(80):
        n = rng.choice(sizes)
        u = rand_units(rng, n, it % 3)
        if it % 4 == 0 and n:
            u[-1] = 0xD800 + rng.randrange(0x400)
        if it % 5 == 0 and n > 10:
            u[7], u[8] = 0xD800 + rng.randrange(0x400), 0xDC00 + rng.randrange(0x400)  # a pair across a vector edge
        n = len(u)
        for be in (False, True):
            src = u.byteswap() if be else u
            want = oracle.to_well_formed_utf16(src, be)
            mis = rng.randrange(8)
            dd = dev(src.view(np.uint8), misalign=2 * mis).view(torch.int16)
            fn = b.to_well_formed_utf16be if be else b.to_well_formed_utf16le
            for omis in (mis, (mis + 3) % 8):
                _, vo = out_buf(n, torch.int16, misalign=omis)
                fn(dd, vo)
                assert vo[:n].cpu().numpy().view(np.uint16).tobytes() == want.tobytes(), (u[:16], n, mis, omis, be)
                check_guard(vo, n)
            val = b.validate_utf16be_with_errors if be else b.validate_utf16le_with_errors
            assert val(vo[:n]) == (0, n)
            fn(dd, dd)  # in place
            assert dd.cpu().numpy().view(np.uint16).tobytes() == want.tobytes()
            if it % 8 == 0:
                h = np.zeros(n + 1, dtype=np.uint16)
                fn(src, h)
                assert h[:n].tobytes() == want.tobytes()
    boms = [b"\xff\xfe", b"\xff\xfe\x00\x00", b"\xfe\xff", b"\x00\x00\xfe\xff", b"\xef\xbb\xbf", b"\xef\xbb", b""]
    for it in range(150):
        n = rng.choice([0, 1, 2, 3, 4, 5, 8, 31, 64, 1000, 4099, 70000])
        text = "".join(rng.choice("aé中😀 ") for _ in range(n))
        d = rng.choice(boms[-2:] if it % 3 else boms) + text.encode(rng.choice(["utf-8", "utf-16-le", "utf-32-le"]))
        if it % 4 == 0:
            d = bytes(rng.randrange(256) for _ in range(n))
        if it % 7 == 0 and d:
            k = rng.randrange(len(d))
            d = d[:k] + bytes([rng.randrange(256)]) + d[k + 1:]
        want = oracle.detect_encodings(d)
        assert b.detect_encodings(dev(d)) == want, (d[:32].hex(), len(d))
        if it % 5 == 0:
            assert b.detect_encodings(d) == want
    from simdutf_b200 import synth
    d = synth.mixed_utf8(1 << 28, seed=4, device="cuda")
    assert b.detect_encodings(d) & 1
    u = synth.mixed_utf16le(1 << 27, seed=5, device="cuda")
    got = b.detect_encodings(u.view(torch.uint8))
    assert got & 2 and not got & 1


def test_binary_to_base64(b, oracle):
    """binary_to_base64 for the four option values (default / url, with and without padding)
    against the oracle, every length class and pointer alignment, device and host path, and a decode round trip."""
    rng = random.Random(64)
    for it in range(120):
        n = rng.choice([0, 1, 2, 3, 4, 5, 47, 48, 49, 95, 96, 97, 1000, 4095, 4096, 4097, 50000, 300001])
        raw = bytes(rng.randrange(256) for _ in range(min(n, 5000))) * (n // 5000 + 1)
        raw = raw[:n]
        for opt in (0, 1, 2, 3):
            want = oracle.binary_to_base64(raw, opt)
            assert b.base64_length_from_binary(n, opt) == len(want)
            for mi, mo in ((0, 0), (rng.randrange(16), rng.randrange(16))):
                d = dev(raw, misalign=mi)
                _, o = out_buf(len(want), torch.uint8, misalign=mo)
                assert b.binary_to_base64(d, o, opt) == len(want), (n, opt, mi, mo)
                assert o[:len(want)].cpu().numpy().tobytes() == want, (n, opt, mi, mo)
                check_guard(o, len(want), 0x5A)
            if it % 10 == 0:
                h = np.zeros(len(want) + 4, dtype=np.uint8)
                assert b.binary_to_base64(raw, h, opt) == len(want) and h[:len(want)].tobytes() == want
    # 64 MiB round trip through the decoder (both alphabets), and the streamed host path (> 48 MiB)
    payload = torch.randint(0, 256, ((1 << 26) + 1,), dtype=torch.uint8, device="cuda")
    for opt in (0, 1):
        nchar = b.base64_length_from_binary(payload.numel(), opt)
        text = torch.empty(nchar, dtype=torch.uint8, device="cuda")
        assert b.binary_to_base64(payload, text, opt) == nchar
        back = torch.empty(payload.numel() + 3, dtype=torch.uint8, device="cuda")
        e, i, k = b.base64_to_binary_details(text, back, opt, 0)
        assert (e, k) == (0, payload.numel()) and torch.equal(back[:k], payload)
    hp = payload.cpu().numpy()
    ht = np.zeros(b.base64_length_from_binary(hp.size, 0) + 4, dtype=np.uint8)
    assert b.binary_to_base64(hp, ht, 0) == ht.size - 4
    assert ht[:ht.size - 4].tobytes() == text_default(b, payload)


def test_base64_from_utf16_input(b, oracle):
    """base64_to_binary for char16_t input: the byte decoder on the narrowed characters, a unit above 0xFF being an
    invalid character (reference src/scalar/base64.h:24-31, :125) — all options, device and host path."""
    rng = random.Random(1616)
    for it in range(80):
        n = rng.choice([0, 1, 3, 4, 5, 63, 64, 65, 1000, 4097, 30000])
        chars = bytearray()
        for _ in range(n):
            x = rng.random()
            chars.append(rng.choice(ABC[:64]) if x < 0.9 else rng.choice(b" \t\n\r\x0c") if x < 0.98 else rng.choice(b"=*"))
        chars += rng.choice([b"", b"=", b"==", b" = "])
        u = np.frombuffer(bytes(chars), dtype=np.uint8).astype(np.uint16)
        if it % 3 == 0 and u.size:
            u[rng.randrange(u.size)] = rng.choice([0x0141, 0x2020, 0xFF41, 0x013D])  # low byte looks valid, unit is not
        narrowed = np.where(u > 0xFF, 0xFF, u).astype(np.uint8).tobytes()
        for opt in ((0, 1, 4, 8) if it % 2 else (0,)):
            for lc in (0, 1, 2):
                want, wout = oracle.base64_to_binary_details(narrowed, opt, lc)
                cap = oracle.maximal_binary_length_from_base64(narrowed)
                d = dev(u.view(np.uint8), dtype=torch.uint8, misalign=2 * rng.randrange(8)).view(torch.int16)
                _, o = out_buf(cap, torch.uint8, misalign=rng.randrange(16))
                got = b.base64_to_binary_details_utf16(d, o, opt, lc)
                assert got[:2] == want[:2], (bytes(chars), opt, lc, got, want)
                if want[0] not in (7, 9):
                    assert got == want and o[:want[2]].cpu().numpy().tobytes() == wout.tobytes()
                if it % 8 == 0:
                    h = np.zeros(cap + 4, dtype=np.uint8)
                    assert b.base64_to_binary_details_utf16(u, h, opt, lc)[:2] == want[:2]


def text_default(b, payload):
    t = torch.empty(b.base64_length_from_binary(payload.numel(), 0), dtype=torch.uint8, device="cuda")
    b.binary_to_base64(payload, t, 0)
    return t.cpu().numpy().tobytes()


def test_repeated_calls_and_epoch_wrap(b, oracle):
    """More than 4096 scan launches on one stream: the 12-bit descriptor epoch wraps and must be handled."""
    from simdutf_b200 import synth
    d = synth.mixed_utf8(70000, seed=51).numpy().tobytes()
    want, o16 = oracle.convert_utf8_to_utf16le_with_errors(d)
    dd = dev(d)
    n16 = want[1]
    _, v = out_buf(n16, torch.int16)
    for it in range(4300):
        assert b.convert_utf8_to_utf16le_with_errors(dd, v) == want
        if it % 500 == 0 or it > 4090:
            assert v[:n16].cpu().numpy().view(np.uint16).tobytes() == o16.tobytes()
            v[:n16].fill_(0)
    check_guard(v, n16, 0x5A5A)


# ------------------------------------------------------------------------------------------------------------
# full-size properties (BASELINE.json configs 1-4) — sizes the oracle cannot walk in seconds
# ------------------------------------------------------------------------------------------------------------
def test_config1_ascii_1gib(b, oracle):
    from simdutf_b200 import synth
    n = 1 << 30
    d = synth.ascii_text(n, seed=1, device="cuda")
    assert b.validate_utf8_with_errors(d) == (0, n)
    assert b.count_utf8(d) == n and b.utf16_length_from_utf8(d) == n
    rng = random.Random(9)
    for pos in (0, 63, 64, 16383, 16384, (1 << 29) + 5, n - 2, n - 1):
        for byte, code in ((0xFF, 1), (0x80, 3), (0xC0, None), (0xE4, 2)):
            old = int(d[pos].item())
            d[pos] = byte
            err, cnt = b.validate_utf8_with_errors(d)
            lo = max(0, pos - 64)
            want = oracle.validate_utf8_with_errors(d[lo:min(n, pos + 64)].cpu().numpy().tobytes())
            assert (err, cnt) == (want[0], want[1] + lo), (pos, byte)
            if code is not None:
                assert err == code and cnt == pos
            d[pos] = old
    assert b.validate_utf8_with_errors(d) == (0, n)


def test_config2_mixed_1gib_roundtrip(b, oracle):
    from simdutf_b200 import synth
    d = synth.mixed_utf8(1 << 30, seed=2, device="cuda")
    n = d.numel()
    assert b.validate_utf8_with_errors(d) == (0, n)
    units = b.utf16_length_from_utf8(d)
    chars = b.count_utf8(d)
    _, out = out_buf(units, torch.int16)
    assert b.convert_utf8_to_utf16le_with_errors(d, out) == (0, units)
    check_guard(out, units, 0x5A5A)
    u16 = out[:units]
    # size-independent properties: counts agree across encodings, and UTF-16 -> UTF-8 gives the input back
    assert b.count_utf16le(u16) == chars
    assert b.utf8_length_from_utf16le(u16) == n
    _, back = out_buf(n, torch.uint8)
    assert b.convert_utf16le_to_utf8_with_errors(u16, back) == (0, n)
    assert torch.equal(back[:n], d)
    # and a slice checked bit-for-bit against the oracle
    head = d[:1 << 22].cpu().numpy().tobytes()
    cut = oracle.trim_partial_utf8(head)
    (e, c), o = oracle.convert_utf8_to_utf16le_with_errors(head[:cut])
    assert e == 0 and u16[:c].cpu().numpy().view(np.uint16).tobytes() == o.tobytes()
    # UTF-32 at a quarter of the size
    qlen = 1 << 28
    while (int(d[qlen].item()) & 0xC0) == 0x80:
        qlen -= 1
    q = d[:qlen]
    nq = b.count_utf8(q)
    _, o32 = out_buf(nq, torch.int32)
    assert b.convert_utf8_to_utf32_with_errors(q, o32) == (0, nq)
    check_guard(o32, nq, 0x5A5A5A5A)
    (e, c), o = oracle.convert_utf8_to_utf32_with_errors(head[:cut])
    assert o32[:c].cpu().numpy().view(np.uint32).tobytes() == o.tobytes()
    # an error deep inside: position must be exact
    p = (1 << 29) + 12345
    while (int(d[p].item()) & 0xC0) == 0x80:
        p -= 1
    d[p] = 0xFF
    assert b.validate_utf8_with_errors(d) == (1, p)
    units_bad = b.utf16_length_from_utf8(d)  # the caller sizes the buffer from the SAME (invalid) input
    del out, back, u16
    _, out = out_buf(units_bad, torch.int16)
    assert b.convert_utf8_to_utf16le_with_errors(d, out) == (1, p)
    check_guard(out, units_bad)


def test_beyond_4gib_offsets(b, oracle):
    """5 GiB of mixed UTF-8 on one GPU (config 5 gives one GPU up to 16 GiB): input offsets beyond 2^32 bytes, more
    than 2^31 output units, tile indices beyond 2^21.  Checked through size-independent properties: counts agree
    across encodings, the round trip gives the input back, the tail agrees with the oracle, an error planted past
    4 GiB is located exactly."""
    from simdutf_b200 import synth
    base = synth.mixed_utf8(1 << 30, seed=9, device="cuda")  # ends on a character boundary
    reps = 5
    d = base.repeat(reps)
    n = d.numel()
    assert n > (1 << 32)
    assert b.validate_utf8_with_errors(d) == (0, n)
    units = b.utf16_length_from_utf8(d)
    assert units == reps * b.utf16_length_from_utf8(base) and units > (1 << 31)
    out = torch.empty(units + 64, dtype=torch.int16, device="cuda")
    out[units:] = 0x5A5A
    assert b.convert_utf8_to_utf16le_with_errors(d, out) == (0, units)
    assert bool((out[units:] == 0x5A5A).all())
    u16 = out[:units]
    per = units // reps
    assert torch.equal(u16[:per], u16[(reps - 1) * per:])  # every repetition transcodes identically
    assert b.utf8_length_from_utf16le(u16) == n and b.count_utf16le(u16) == reps * b.count_utf8(base)
    back = torch.empty(n, dtype=torch.uint8, device="cuda")
    assert b.convert_utf16le_to_utf8_with_errors(u16, back) == (0, n)
    assert torch.equal(back, d)
    tail = base[-(1 << 20):].cpu().numpy().tobytes()
    k = 0
    while (tail[k] & 0xC0) == 0x80:
        k += 1
    (e, c), o = oracle.convert_utf8_to_utf16le_with_errors(tail[k:])
    assert e == 0 and u16[units - c:].cpu().numpy().view(np.uint16).tobytes() == o.tobytes()
    del back
    p = (1 << 32) + (1 << 29) + 777
    while (int(d[p].item()) & 0xC0) == 0x80:
        p -= 1
    d[p] = 0xF8
    assert b.validate_utf8_with_errors(d) == (1, p)
    assert b.convert_utf8_to_utf16le_with_errors(d, out)[:2] == (1, p)


def test_config3_utf16_2gib(b, oracle):
    from simdutf_b200 import synth
    u = synth.mixed_utf16le(1 << 30, seed=3, device="cuda")
    n = u.numel()
    nbytes = b.utf8_length_from_utf16le(u)
    chars = b.count_utf16le(u)
    assert b.validate_utf16le_with_errors(u) == (0, n)
    _, out = out_buf(nbytes, torch.uint8)
    assert b.convert_utf16le_to_utf8_with_errors(u, out) == (0, nbytes)
    check_guard(out, nbytes)
    o8 = out[:nbytes]
    assert b.validate_utf8_with_errors(o8) == (0, nbytes)
    assert b.count_utf8(o8) == chars and b.utf16_length_from_utf8(o8) == n
    head = u[:1 << 21].cpu().numpy().view(np.uint16)
    if (head[-1] & 0xFC00) == 0xD800:
        head = head[:-1]
    (e, c), o = oracle.convert_utf16le_to_utf8_with_errors(head)
    assert e == 0 and o8[:c].cpu().numpy().tobytes() == o.tobytes()
    # back to UTF-16: identical units
    _, back = out_buf(n, torch.int16)
    assert b.convert_utf8_to_utf16le_with_errors(o8, back) == (0, n)
    assert torch.equal(back[:n], u)
    # one injected unpaired surrogate at 0.9*len (BASELINE.json config 3)
    p = int(0.9 * n)
    while (int(u[p].item()) & 0xF800) == 0xD800:
        p += 1
    if (int(u[p - 1].item()) & 0xFC00) == 0xD800:  # previous is a high surrogate: it becomes the error instead
        p += 1
    old = int(u[p].item())
    u[p] = -9216  # 0xDC00
    assert b.convert_utf16le_to_utf8_with_errors(u, out) == (6, p)
    assert b.validate_utf16le_with_errors(u) == (6, p)
    u[p] = -10240  # 0xD800 followed by a non-surrogate
    assert b.convert_utf16le_to_utf8_with_errors(u, out) == (6, p)
    u[p] = old


def test_config4_base64_2gib(b, oracle):
    from simdutf_b200 import synth
    for url in (False, True):
        text, payload = synth.base64_text(1 << 31, seed=4, device="cuda", url=url)
        opt = 1 if url else 0
        cap = text.numel() // 4 * 3 + 3
        _, out = out_buf(cap, torch.uint8)
        e, i, n = b.base64_to_binary_details(text, out, opt, 0)
        assert (e, n) == (0, payload.numel()), (e, i, n, payload.numel())
        assert torch.equal(out[:n], payload)
        check_guard(out, cap)
        q = (1 << 30) + 77
        old = int(text[q].item())
        text[q] = ord("=")
        assert b.base64_to_binary(text, out, opt, 0) == (7, q)
        text[q] = 0xC3
        assert b.base64_to_binary(text, out, opt, 0) == (7, q)
        text[q] = old
        del text, payload, out
        torch.cuda.empty_cache()


# ------------------------------------------------------------------------------------------------------------
# streaming host path (capi.cu run_host_streamed): buffers of several 32 MiB segments, host pointers
# ------------------------------------------------------------------------------------------------------------
def test_host_streaming_segments(b, oracle):
    from simdutf_b200 import synth
    data = synth.mixed_utf8(150 << 20, seed=11).numpy()
    host = data.tobytes()
    n16, n8 = oracle.utf16_length_from_utf8(host), oracle.count_utf8(host)
    assert b.validate_utf8_with_errors(data) == (0, len(host))
    assert b.count_utf8(data) == n8 and b.utf16_length_from_utf8(data) == n16
    want16, o16 = oracle.convert_utf8_to_utf16le_with_errors(host)
    h16 = np.full(n16 + GUARD, 0x5A5A, dtype=np.uint16)
    assert b.convert_utf8_to_utf16le_with_errors(data, h16) == want16 == (0, n16)
    assert (h16[n16:] == 0x5A5A).all() and h16[:n16].tobytes() == o16.tobytes()
    want32, o32 = oracle.convert_utf8_to_utf32_with_errors(host)
    h32 = np.full(n8 + GUARD, 0x5A5A5A5A, dtype=np.uint32)
    assert b.convert_utf8_to_utf32_with_errors(data, h32) == want32
    assert h32[:n8].tobytes() == o32.tobytes()
    # UTF-16LE -> UTF-8 brings the input back (2 x 150 MiB-ish of units: ten segments)
    back = np.full(len(host) + GUARD, 0x5A, dtype=np.uint8)
    assert b.validate_utf16le_with_errors(h16[:n16]) == (0, n16)
    assert b.count_utf16le(h16[:n16]) == n8 and b.utf8_length_from_utf16le(h16[:n16]) == len(host)
    assert b.convert_utf16le_to_utf8_with_errors(h16[:n16], back) == (0, len(host))
    assert back[:len(host)].tobytes() == host and (back[len(host):] == 0x5A).all()
    # one bad byte in a late segment, one right at a segment seam, one in the first segment
    for pos in (len(host) - 12345, (64 << 20) + 1, 7):
        bad = data.copy()
        bad[pos] = 0xFF
        want = oracle.validate_utf8_with_errors(bad.tobytes())
        assert b.validate_utf8_with_errors(bad) == want
        assert b.convert_utf8_to_utf16le_with_errors(bad, h16) == want
    # a lone surrogate in a late segment of the UTF-16 input
    u = h16[:n16].copy()
    p = n16 - 4321
    while (u[p] & 0xF800) == 0xD800 or (u[p - 1] & 0xFC00) == 0xD800:
        p -= 1
    u[p] = 0xDC00
    assert b.validate_utf16le_with_errors(u) == (6, p)
    assert b.convert_utf16le_to_utf8_with_errors(u, back) == (6, p)


def test_batch_of_small_strings(b, oracle):
    """many small strings per launch.  Every string of a batch gets the result the
    single-string entry point gives (the oracle's), for valid, truncated, overlong, surrogate and too-large
    strings, empty strings included, in one launch per operation (reference
    tests/validate_utf8_with_errors_tests.cpp:54-69 walks such strings one call at a time)."""
    import random
    rng = random.Random(20261018)
    pool = ["a", "é", "€", "😀", "\x00", "z" * 40, "日本語" * 11, "😀" * 9]
    strings = [b"", b"A", b"\xff", b"\x80", b"\xc0\x80", b"\xed\xa0\x80", b"\xf4\x90\x80\x80", b"\xe2\x82", b"ab\xf0\x9f\x98"]
    for _ in range(3000):
        k = rng.randrange(0, 60)
        s = "".join(rng.choice(pool) for _ in range(k)).encode()
        roll = rng.random()
        if roll < 0.15 and s:      # cut a character short / corrupt a byte
            s = s[: rng.randrange(1, len(s) + 1)]
        elif roll < 0.25 and s:
            i = rng.randrange(len(s))
            s = s[:i] + bytes([rng.choice([0x80, 0xC0, 0xF5, 0xFF, 0xED, 0xA0])]) + s[i + 1:]
        strings.append(s)
    strings += [("x" * n).encode() for n in (31, 32, 33, 1023, 1024, 1025)] + [("é" * 700).encode(), ("€" * 341 + "a").encode()]
    launches0 = b.launch_count()
    got_v = b.validate_utf8_batch(strings)
    got_l = b.utf16_length_from_utf8_batch(strings)
    got_c = b.count_utf8_batch(strings)
    got_t = b.convert_utf8_to_utf16le_batch(strings)
    assert b.launch_count() - launches0 == 4, "a batch is ONE launch per operation"
    for i, s in enumerate(strings):
        assert got_v[i] == oracle.validate_utf8_with_errors(s), (i, s)
        assert got_l[i] == oracle.utf16_length_from_utf8(s), (i, s)
        assert got_c[i] == oracle.count_utf8(s), (i, s)
        (werr, wcnt), wout = oracle.convert_utf8_to_utf16le_with_errors(s)
        assert got_t[i][0] == (werr, wcnt), (i, s)
        if werr == 0:
            assert got_t[i][1] == wout.tobytes(), (i, s)
    # device flavour on a packed buffer at odd offsets
    import torch
    packed = b"".join(strings)
    offs = [0]
    for s in strings:
        offs.append(offs[-1] + len(s))
    d = torch.frombuffer(bytearray(b"#" + packed), dtype=torch.uint8).cuda()[1:]
    o = torch.tensor(offs, dtype=torch.int64).cuda()
    r = b.validate_utf8_batch_device(d, o).cpu()
    for i, s in enumerate(strings):
        assert (int(r[i, 0]) & 0xFFFFFFFF, int(r[i, 1])) == oracle.validate_utf8_with_errors(s), (i, s)
    assert b.validate_utf8_batch([]) == []
