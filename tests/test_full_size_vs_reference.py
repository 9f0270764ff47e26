"""GPU tests (-m gpu): BASELINE.json configs 1-4 at their FULL sizes, every output byte and every result compared with
the UNMODIFIED reference (oracle/_ref/libsimdutf_ref.so, its best kernel on the box: icelake / haswell) — SURVEY.md
§8d "output compared byte-for-byte with icelake".  The reference walks 1-2 GiB in about a second, so nothing here is
sampled: the whole 1 GiB / 2 GiB outputs are brought back and compared.

Nothing reads /root/reference at run time; the prebuilt reference library travels with the snapshot.
"""
import ctypes

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def b():
    import simdutf_b200
    simdutf_b200.load()
    assert simdutf_b200.device_count() >= 1, "no sm_100 device visible: the CUDA path cannot run"
    simdutf_b200.set_device(0)
    return simdutf_b200


@pytest.fixture(scope="module")
def R(ref):
    if ref is None:
        pytest.skip("oracle/_ref/libsimdutf_ref.so did not travel: full-size byte-for-byte comparison impossible")
    return ref


class _Res(ctypes.Structure):
    _fields_ = [("error", ctypes.c_int32), ("count", ctypes.c_uint64)]


class _Full(ctypes.Structure):
    _fields_ = [("error", ctypes.c_int32), ("input_count", ctypes.c_uint64), ("output_count", ctypes.c_uint64)]


def _vp(a):
    return ctypes.c_void_p(a.ctypes.data)


def _ref_validate(R, host):
    r = _Res()
    assert R.L.ref_validate_utf8_with_errors(R.best.encode(), _vp(host), ctypes.c_size_t(host.size), ctypes.byref(r)) == 0
    return (r.error, r.count)


def test_config1_validate_ascii_1gib_vs_reference(b, R):
    from simdutf_b200 import synth
    n = 1 << 30
    d = synth.ascii_text(n, seed=1, device="cuda")
    host = d.cpu().numpy()
    assert b.validate_utf8_with_errors(d) == _ref_validate(R, host) == (0, n)
    assert b.count_utf8(d) == int(R.L.ref_count_utf8(R.best.encode(), _vp(host), ctypes.c_size_t(n))) == n
    # SURVEY §8d cfg 1 add-on: one corrupted byte at several positions, each error class, the reference's verdict on the
    # WHOLE buffer as the expectation
    for pos in (0, 63, 64, 16383, 16384, (1 << 29) + 5, n - 2, n - 1):
        for byte in (0xFF, 0x80, 0xC0, 0xE4, 0xF5, 0xED):
            old = int(host[pos])
            host[pos] = byte
            d[pos] = byte
            assert b.validate_utf8_with_errors(d) == _ref_validate(R, host), (pos, hex(byte))
            host[pos] = old
            d[pos] = old


def test_config2_convert_mixed_1gib_vs_reference(b, R):
    from simdutf_b200 import synth
    d = synth.mixed_utf8(1 << 30, seed=2, device="cuda")
    n = d.numel()
    host = d.cpu().numpy()
    units = b.utf16_length_from_utf8(d)
    assert units == int(R.L.ref_utf16_length_from_utf8(R.best.encode(), _vp(host), ctypes.c_size_t(n)))
    out = torch.full((units + 64,), 0x5A5A, dtype=torch.int16, device="cuda")
    assert b.convert_utf8_to_utf16le_with_errors(d, out) == (0, units)
    assert bool((out[units:] == 0x5A5A).all()), "output buffer overrun"
    want = np.empty(units + 64, dtype=np.uint16)
    r = _Res()
    assert R.L.ref_convert_utf8_to_utf16le_with_errors(R.best.encode(), _vp(host), ctypes.c_size_t(n), _vp(want), ctypes.byref(r)) == 0
    assert (r.error, r.count) == (0, units)
    got = out[:units].cpu().numpy().view(np.uint16)
    assert np.array_equal(got, want[:units]), "UTF-16LE output differs from the reference"
    del out, got
    # UTF-32 of the same buffer, whole output
    chars = b.count_utf8(d)
    o32 = torch.full((chars + 64,), 0x5A5A5A5A, dtype=torch.int32, device="cuda")
    assert b.convert_utf8_to_utf32_with_errors(d, o32) == (0, chars)
    assert bool((o32[chars:] == 0x5A5A5A5A).all())
    want32 = np.empty(chars + 64, dtype=np.uint32)
    assert R.L.ref_convert_utf8_to_utf32_with_errors(R.best.encode(), _vp(host), ctypes.c_size_t(n), _vp(want32), ctypes.byref(r)) == 0
    assert (r.error, r.count) == (0, chars)
    assert np.array_equal(o32[:chars].cpu().numpy().view(np.uint32), want32[:chars]), "UTF-32 output differs from the reference"
    del o32, want32
    # an error deep inside: (error, position) of the reference on the whole buffer
    p = (1 << 29) + 12345
    while (int(host[p]) & 0xC0) == 0x80:
        p -= 1
    for byte in (0xFF, 0xC0, 0x80):
        old = int(host[p])
        host[p] = byte
        d[p] = byte
        assert R.L.ref_convert_utf8_to_utf16le_with_errors(R.best.encode(), _vp(host), ctypes.c_size_t(n), _vp(want), ctypes.byref(r)) == 0
        ub = b.utf16_length_from_utf8(d)
        o = torch.empty(ub + 64, dtype=torch.int16, device="cuda")
        assert b.convert_utf8_to_utf16le_with_errors(d, o) == (r.error, r.count) == (r.error, p), (hex(byte), r.error, r.count)
        host[p] = old
        d[p] = old
        del o


def test_config3_utf16_2gib_vs_reference(b, R):
    from simdutf_b200 import synth
    u = synth.mixed_utf16le(1 << 30, seed=3, device="cuda")
    n = u.numel()
    host = u.cpu().numpy().view(np.uint16)
    nbytes = b.utf8_length_from_utf16le(u)
    assert nbytes == int(R.L.ref_utf8_length_from_utf16le(R.best.encode(), _vp(host), ctypes.c_size_t(n)))
    assert b.count_utf16le(u) == int(R.L.ref_count_utf16le(R.best.encode(), _vp(host), ctypes.c_size_t(n)))
    out = torch.full((nbytes + 64,), 0x5A, dtype=torch.uint8, device="cuda")
    assert b.convert_utf16le_to_utf8_with_errors(u, out) == (0, nbytes)
    assert bool((out[nbytes:] == 0x5A).all()), "output buffer overrun"
    want = np.empty(nbytes + 64, dtype=np.uint8)
    r = _Res()
    assert R.L.ref_convert_utf16le_to_utf8_with_errors(R.best.encode(), _vp(host), ctypes.c_size_t(n), _vp(want), ctypes.byref(r)) == 0
    assert (r.error, r.count) == (0, nbytes)
    assert np.array_equal(out[:nbytes].cpu().numpy(), want[:nbytes]), "UTF-8 output differs from the reference"
    # run B: one injected unpaired surrogate at 0.9 * len, both variants
    p = int(0.9 * n)
    while (int(host[p]) & 0xF800) == 0xD800:
        p += 1
    if (int(host[p - 1]) & 0xFC00) == 0xD800:
        p += 1
    old = int(host[p])
    for unit in (0xDC00, 0xD800):
        host[p] = unit
        u[p] = unit - 0x10000
        assert R.L.ref_convert_utf16le_to_utf8_with_errors(R.best.encode(), _vp(host), ctypes.c_size_t(n), _vp(want), ctypes.byref(r)) == 0
        assert (r.error, r.count) == (6, p)
        assert b.convert_utf16le_to_utf8_with_errors(u, out) == (6, p)
        assert b.validate_utf16le_with_errors(u) == (6, p)
        # the output prefix for units < p is the reference's prefix
        pre = b.utf8_length_from_utf16le(u[:p])
        assert np.array_equal(out[:pre].cpu().numpy(), want[:pre])
    host[p] = old


@pytest.mark.parametrize("url", [False, True])
def test_config4_base64_2gib_vs_reference(b, R, url):
    from simdutf_b200 import synth
    text, payload = synth.base64_text(1 << 31, seed=4, device="cuda", url=url)
    opt = 1 if url else 0
    nt = text.numel()
    host = text.cpu().numpy()
    cap = nt // 4 * 3 + 3
    out = torch.full((cap + 64,), 0x5A, dtype=torch.uint8, device="cuda")
    got = b.base64_to_binary_details(text, out, opt, 0)
    want = np.empty(cap + 64, dtype=np.uint8)
    f = _Full()
    assert R.L.ref_base64_to_binary_details(R.best.encode(), _vp(host), ctypes.c_size_t(nt), _vp(want), ctypes.c_uint64(opt),
                                            ctypes.c_uint64(0), ctypes.byref(f)) == 0
    assert got == (f.error, f.input_count, f.output_count) and f.error == 0 and f.output_count == payload.numel()
    assert bool((out[cap:] == 0x5A).all()), "output buffer overrun"
    assert np.array_equal(out[:f.output_count].cpu().numpy(), want[:f.output_count]), "decoded bytes differ from the reference"
    # parity add-ons: an invalid character mid-stream, at the reference's position
    q = (1 << 30) + 77
    for ch in (ord("="), 0xC3, ord("*")):
        old = int(host[q])
        host[q] = ch
        text[q] = ch
        r = _Res()
        assert R.L.ref_base64_to_binary(R.best.encode(), _vp(host), ctypes.c_size_t(nt), _vp(want), ctypes.c_uint64(opt), ctypes.c_uint64(0),
                                        ctypes.byref(r)) == 0
        assert b.base64_to_binary(text, out, opt, 0) == (r.error, r.count) == (7, q)
        host[q] = old
        text[q] = old
