"""Seeded synthetic inputs of SURVEY.md §8(d) / BASELINE.json configs, built with torch ops so the same
generator runs on the GPU (bench, full-size parity properties) and on the CPU (oracle-sized parity tests,
CPU baseline sample).  Generation is plumbing, not the product: nothing here is on the timed path.

All generators return a 1-D uint8 tensor (UTF-8 / base64 text) or int16 tensor viewed as UTF-16LE units.
They work in bounded pieces so that a 16 GiB shard never needs more than a few GiB of temporaries.
"""
from __future__ import annotations

import torch

_PIECE = 1 << 25  # code points (or bytes) generated per piece


def _gen(seed: int, device) -> torch.Generator:
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    return g


def ascii_text(nbytes: int, seed: int = 1, device="cpu") -> torch.Tensor:
    """Config 1: bytes uniform in 0x20..0x7E with '\\n' every 80th byte."""
    out = torch.empty(nbytes, dtype=torch.uint8, device=device)
    g = _gen(seed, device)
    for lo in range(0, nbytes, _PIECE * 4):
        hi = min(nbytes, lo + _PIECE * 4)
        out[lo:hi] = torch.randint(0x20, 0x7F, (hi - lo,), generator=g, device=device, dtype=torch.uint8)
    out[79::80] = 0x0A
    return out


def _mixed_code_points(n: int, g: torch.Generator, device, classes) -> tuple[torch.Tensor, torch.Tensor]:
    """n code points, class i.i.d. uniform over `classes` = [(lo, hi_exclusive), ...]; returns (cp, class)."""
    cls = torch.randint(0, len(classes), (n,), generator=g, device=device)
    r = torch.rand(n, generator=g, device=device, dtype=torch.float64)
    cp = torch.zeros(n, dtype=torch.int64, device=device)
    for i, (lo, hi) in enumerate(classes):
        m = cls == i
        cp = torch.where(m, (lo + (r * (hi - lo)).floor()).to(torch.int64), cp)
    return cp, cls


UTF8_MIX = [(0x20, 0x7F), (0xA0, 0x250), (0x4E00, 0xA000), (0x1F300, 0x1F650)]  # ASCII / Latin / CJK / emoji


def encode_utf8(cp: torch.Tensor) -> torch.Tensor:
    """Vectorised UTF-8 encoder for valid scalar values (used only to BUILD inputs)."""
    n1 = cp < 0x80
    n2 = (cp >= 0x80) & (cp < 0x800)
    n3 = (cp >= 0x800) & (cp < 0x10000)
    lens = torch.where(n1, 1, torch.where(n2, 2, torch.where(n3, 3, 4)))
    off = torch.cumsum(lens, 0) - lens
    total = int(lens.sum().item()) if cp.numel() else 0
    out = torch.zeros(total, dtype=torch.uint8, device=cp.device)
    b0 = torch.where(n1, cp, torch.where(n2, 0xC0 | (cp >> 6), torch.where(n3, 0xE0 | (cp >> 12), 0xF0 | (cp >> 18))))
    out[off] = b0.to(torch.uint8)
    # continuation bytes, counted from the END of each character
    last = 0x80 | (cp & 0x3F)
    m = lens >= 2
    out[(off + lens - 1)[m]] = last[m].to(torch.uint8)
    m = lens >= 3
    out[(off + lens - 2)[m]] = (0x80 | ((cp >> 6) & 0x3F))[m].to(torch.uint8)
    m = lens == 4
    out[(off + 1)[m]] = (0x80 | ((cp >> 12) & 0x3F))[m].to(torch.uint8)
    return out


def mixed_utf8(nbytes: int, seed: int = 2, device="cpu", classes=UTF8_MIX) -> torch.Tensor:
    """Config 2/5: valid UTF-8, code points i.i.d. 25% each of 1/2/3/4-byte classes; the result is the
    longest whole-character prefix that fits in `nbytes` (so len <= nbytes, ends on a character boundary)."""
    g = _gen(seed, device)
    pieces, have = [], 0
    while have < nbytes:
        want = nbytes - have
        n = max(16, min(_PIECE, int(want / 2.4) + 16))
        cp, _ = _mixed_code_points(n, g, device, classes)
        b = encode_utf8(cp)
        if b.numel() > want:  # cut at a character boundary
            cut = want
            while cut > 0 and (int(b[cut].item()) & 0xC0) == 0x80:
                cut -= 1
            b = b[:cut]
            pieces.append(b)
            have += b.numel()
            break
        pieces.append(b)
        have += b.numel()
    return torch.cat(pieces) if len(pieces) != 1 else pieces[0]


# ------------------------------------------------------------------------------------------------------------
# Config 5: ONE global buffer of any size, addressable by byte range, so that every rank of a sharded run can
# materialise exactly its own slice [c_k, c_k+1) of it (plus a few bytes around the nominal cut points) without any
# rank ever holding the whole 16 GiB.  The stream is a sequence of independent fixed-size blocks, block k generated
# by a counter-based seed (seed, k); a block is whole characters of the config-2 distribution padded with <= 3
# ASCII '.' to exactly STREAM_BLOCK bytes.  STREAM_BLOCK is odd, so the nominal cuts k*N/G of a power-of-two total
# never coincide with block boundaries: they land mid-distribution, inside a character 60 % of the time.
# ------------------------------------------------------------------------------------------------------------
STREAM_BLOCK = (1 << 26) + 1


def stream_block(seed: int, k: int, device="cpu", block: int = STREAM_BLOCK) -> torch.Tensor:
    b = mixed_utf8(block, seed=(seed * 1000003 + k * 7919 + 12345) & 0x7FFFFFFF, device=device)
    pad = block - int(b.numel())
    if pad:
        b = torch.cat([b, torch.full((pad,), 0x2E, dtype=torch.uint8, device=device)])
    return b


def stream_range(seed: int, lo: int, hi: int, device="cpu", block: int = STREAM_BLOCK) -> torch.Tensor:
    """Bytes [lo, hi) of the global stream."""
    out = torch.empty(max(0, hi - lo), dtype=torch.uint8, device=device)
    k = lo // block
    pos = lo
    while pos < hi:
        b = stream_block(seed, k, device, block)
        a = pos - k * block
        take = min(block - a, hi - pos)
        out[pos - lo: pos - lo + take] = b[a: a + take]
        pos += take
        k += 1
    return out


def stream_total_len(seed: int, nominal: int, device="cpu", block: int = STREAM_BLOCK) -> int:
    """Length of the longest whole-character prefix of the stream that fits in `nominal` bytes."""
    lo = max(0, nominal - 4)
    tail = stream_range(seed, lo, nominal + 1, device, block).cpu().tolist()
    cut = nominal
    for _ in range(3):
        if cut <= 0 or (tail[cut - lo] & 0xC0) != 0x80:
            break
        cut -= 1
    return cut


UTF16_MIX = [(0x20, 0x7F), (0xA0, 0x800), (0x800, 0xD800), (0x10000, 0x110000)]


def mixed_utf16le(nunits: int, seed: int = 3, device="cpu", classes=UTF16_MIX) -> torch.Tensor:
    """Config 3: valid UTF-16LE as an int16 tensor of `<= nunits` units (whole characters only)."""
    g = _gen(seed, device)
    pieces, have = [], 0
    while have < nunits:
        want = nunits - have
        n = max(16, min(_PIECE, int(want / 1.2) + 16))
        cp, _ = _mixed_code_points(n, g, device, classes)
        sup = cp >= 0x10000
        lens = torch.where(sup, 2, 1)
        off = torch.cumsum(lens, 0) - lens
        total = int(lens.sum().item())
        u = torch.zeros(total, dtype=torch.int64, device=device)
        v = cp - 0x10000
        u[off] = torch.where(sup, 0xD800 + (v >> 10), cp)
        u[(off + 1)[sup]] = (0xDC00 + (v & 0x3FF))[sup]
        if total > want:
            cut = want
            if cut > 0 and (int(u[cut].item()) & 0xFC00) == 0xDC00:
                cut -= 1
            u = u[:cut]
            pieces.append(_to_int16(u))
            have += u.numel()
            break
        pieces.append(_to_int16(u))
        have += total
    return torch.cat(pieces) if len(pieces) != 1 else pieces[0]


def _to_int16(u: torch.Tensor) -> torch.Tensor:
    # values 0..65535 -> int16 bit pattern
    return torch.where(u >= 0x8000, u - 0x10000, u).to(torch.int16)


_STD = "ABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz0123456789+/"
_URL = "ABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz0123456789-_"


def _b64_encode(payload: torch.Tensor, alphabet: torch.Tensor, pad: bool) -> torch.Tensor:
    """Plain base64 of `payload` (no line breaks)."""
    device = payload.device
    nbytes = payload.numel()
    full = nbytes // 3
    p = payload[: full * 3].to(torch.int32).view(-1, 3)
    t = (p[:, 0] << 16) | (p[:, 1] << 8) | p[:, 2]
    sx = torch.stack([(t >> 18) & 63, (t >> 12) & 63, (t >> 6) & 63, t & 63], dim=1).reshape(-1)
    tail_bytes = nbytes - full * 3
    tail = []
    if tail_bytes == 1:
        b = int(payload[-1].item())
        tail = [b >> 2, (b & 3) << 4]
    elif tail_bytes == 2:
        b0, b1 = int(payload[-2].item()), int(payload[-1].item())
        tail = [b0 >> 2, ((b0 & 3) << 4) | (b1 >> 4), (b1 & 15) << 2]
    if tail:
        sx = torch.cat([sx, torch.tensor(tail, dtype=sx.dtype, device=device)])
    chars = alphabet[sx.long()]
    if tail and pad:
        chars = torch.cat([chars, torch.full((4 - len(tail),), ord("="), dtype=torch.uint8, device=device)])
    return chars


def _wrap_lines(chars: torch.Tensor, line: int) -> torch.Tensor:
    """CRLF after every `line` characters (a trailing partial line gets none)."""
    n = chars.numel()
    whole = n // line
    body = chars[: whole * line].view(whole, line)
    crlf = torch.tensor([13, 10], dtype=torch.uint8, device=chars.device).expand(whole, 2)
    return torch.cat([torch.cat([body, crlf], dim=1).reshape(-1), chars[whole * line:]])


def _sprinkle_ws(text: torch.Tensor, prob: float, g: torch.Generator) -> torch.Tensor:
    """Insert ' ' or '\\t' before a random `prob` fraction of the characters."""
    n = text.numel()
    if prob <= 0 or n == 0:
        return text
    device = text.device
    ins = torch.rand(n, generator=g, device=device) < prob
    dst = torch.arange(n, device=device) + torch.cumsum(ins, 0)
    total = int(dst[-1].item()) + 1
    out = torch.where(torch.rand(total, generator=g, device=device) < 0.5, 0x20, 0x09).to(torch.uint8)
    out[dst] = text
    return out


def base64_text(nchars: int, seed: int = 4, device="cpu", url: bool = False, line: int = 76,
                sparse_ws: float = 0.001) -> tuple[torch.Tensor, torch.Tensor]:
    """Config 4: about `nchars` characters of base64 text: random bytes encoded with the standard (padded) or
    URL (unpadded) alphabet, CRLF after every `line` characters, plus sparse ' ' / '\\t' (probability
    `sparse_ws` per character).  The payload length is 3k+1, so the standard variant ends in "==".
    Returns (text uint8, the binary payload uint8 it decodes to).  `line` must be a multiple of 4."""
    assert line % 4 == 0
    g = _gen(seed, device)
    alphabet = torch.tensor(list((_URL if url else _STD).encode()), dtype=torch.uint8, device=device)
    chars_per_line = line + 2
    nlines = max(1, int(nchars / (1.0 + sparse_ws)) // chars_per_line)
    lines_per_piece = max(1, (1 << 27) // chars_per_line)
    texts, payloads = [], []
    done = 0
    while done < nlines:
        take = min(lines_per_piece, nlines - done)
        last = done + take == nlines
        nbytes = take * line // 4 * 3
        if last:
            nbytes = max(1, nbytes - 2)  # 3k+1 bytes: the stream ends with a 2-sextet group (+ "==")
        payload = torch.randint(0, 256, (nbytes,), generator=g, device=device, dtype=torch.uint8)
        chars = _b64_encode(payload, alphabet, pad=not url)
        texts.append(_sprinkle_ws(_wrap_lines(chars, line), sparse_ws, g))
        payloads.append(payload)
        done += take
    text = torch.cat(texts) if len(texts) != 1 else texts[0]
    payload = torch.cat(payloads) if len(payloads) != 1 else payloads[0]
    return text, payload
