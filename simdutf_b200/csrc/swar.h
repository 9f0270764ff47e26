// swar.h — byte-level (SWAR) helpers shared by the sm_100a kernels: the count predicates of the reduction passes, the
// exact per-position verdicts that pin down an error once the bit-plane detectors (bitplane.h) have flagged a block,
// the base64 class function and the last-chunk rules.
//
// Everything here is `__host__ __device__` and free of CUDA-only constructs so
// that the exact same code is exercised on the CPU by tests/host/ (there is no
// GPU in the build container) and on the GPU by the kernels.  It is NOT a CPU
// fallback: nothing in the product path calls these functions on the host.
//
// Conventions
//   * A "granule" is 16 input bytes at a 16-byte-aligned address, held as four
//     little-endian 32-bit words w[0..3]; byte k of the granule is bits
//     [8*(k&3), 8*(k&3)+8) of w[k>>2].
//   * Predicates over bytes are "bit-7 masks": a 32-bit word whose bit 7 of
//     byte k is the predicate for byte k (all other bits zero).
//   * `prev` / `next` are the 32-bit words immediately before / after the
//     granule in the byte stream (zero filler outside the buffer).
//
// Semantics implemented (reference file:line in each function's comment) are
// those of simdutf's scalar routines, the ground truth for every
// (error, position) pair the reference's SIMD kernels return.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define B200_HD __host__ __device__ __forceinline__
#else
#define B200_HD inline
#endif

namespace b200 {

// simdutf::error_code (reference include/simdutf/error.h:5-32)
enum : int {
  kSuccess = 0, kHeaderBits = 1, kTooShort = 2, kTooLong = 3, kOverlong = 4, kTooLarge = 5,
  kSurrogate = 6, kInvalidBase64Character = 7, kBase64InputRemainder = 8, kBase64ExtraBits = 9,
  kOutputBufferTooSmall = 10, kOther = 11
};

constexpr uint32_t kH = 0x80808080u;

// PRMT: result byte i = byte (sel>>(4i) & 7) of the 8-byte pool {a (bytes 0-3), b (bytes 4-7)}.
B200_HD uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) {
#if defined(__CUDA_ARCH__)
  return __byte_perm(a, b, sel);
#else
  uint64_t pool = (uint64_t)a | ((uint64_t)b << 32);
  uint32_t r = 0;
  for (int i = 0; i < 4; i++) r |= (uint32_t)((pool >> (8 * ((sel >> (4 * i)) & 7))) & 0xFF) << (8 * i);
  return r;
#endif
}
B200_HD int popc(uint32_t x) {
#if defined(__CUDA_ARCH__)
  return __popc(x);
#else
  return __builtin_popcount(x);
#endif
}

// Stream shifts.  backK(prev,cur)[j] = stream byte (j-K) relative to cur; fwdK(cur,next)[j] = stream byte (j+K).
B200_HD uint32_t back2(uint32_t prev, uint32_t cur) { return prmt(prev, cur, 0x5432); }
B200_HD uint32_t fwd1(uint32_t cur, uint32_t next) { return prmt(cur, next, 0x4321); }

// Expand 4 bits into a bit-7 mask word.
B200_HD uint32_t unmask4(uint32_t b) {
  return ((b & 1) << 7) | ((b & 2) << 14) | ((b & 4) << 21) | ((b & 8) << 28);
}

// ---------------------------------------------------------------------------------------------
// UTF-8 byte predicates (bit-7 masks) used by the counting passes
// ---------------------------------------------------------------------------------------------
// Non-continuation bytes (code-point starts), reference src/scalar/utf8.h:230-241: (int8_t)b > -65.
B200_HD uint32_t u8_noncont(uint32_t w) { return (~w | (w << 1)) & kH; }
// Bytes >= 0xF0, reference src/scalar/utf8.h:243-255 (the second unit of a surrogate pair).
B200_HD uint32_t u8_ge_f0(uint32_t w) { return w & (w << 1) & (w << 2) & (w << 3) & kH; }

// True iff the last three bytes of the buffer start a sequence that the buffer truncates.
B200_HD bool u8_incomplete_tail(uint32_t b_last, uint32_t b_last2, uint32_t b_last3) {
  return b_last >= 0xC0 || b_last2 >= 0xE0 || b_last3 >= 0xF0;
}

// Exact local verdict for byte position i of a buffer of `len` bytes: the error the
// scalar parser (reference src/scalar/utf8.h:102-200) reports AT i if it reaches i as a character
// start, or TOO_LONG for a continuation byte no preceding lead accounts for.  The first error of the
// buffer is the minimum-position non-zero verdict.  `at(j)` returns byte j (0 <= j < len).
template <class At>
B200_HD int u8_verdict(At &&at, uint64_t i, uint64_t len) {
  const uint32_t b0 = at(i);
  if (b0 < 0x80) return kSuccess;
  if ((b0 & 0xC0) == 0x80) {
    for (uint32_t k = 1; k <= 3; k++) {
      if (i < k) return kTooLong;
      const uint32_t b = at(i - k);
      if ((b & 0xC0) != 0x80) {
        const uint32_t L = b < 0xC0 ? 1u : b < 0xE0 ? 2u : b < 0xF0 ? 3u : b < 0xF8 ? 4u : 1u;
        return L > k ? kSuccess : kTooLong;
      }
    }
    return kTooLong;
  }
  if ((b0 & 0xE0) == 0xC0) {
    if (i + 2 > len) return kTooShort;
    const uint32_t b1 = at(i + 1);
    if ((b1 & 0xC0) != 0x80) return kTooShort;
    return (b0 & 0x1E) == 0 ? kOverlong : kSuccess;
  }
  if ((b0 & 0xF0) == 0xE0) {
    if (i + 3 > len) return kTooShort;
    const uint32_t b1 = at(i + 1), b2 = at(i + 2);
    if ((b1 & 0xC0) != 0x80 || (b2 & 0xC0) != 0x80) return kTooShort;
    const uint32_t cp = ((b0 & 0x0F) << 12) | ((b1 & 0x3F) << 6) | (b2 & 0x3F);
    if (cp < 0x800) return kOverlong;
    if (cp >= 0xD800 && cp <= 0xDFFF) return kSurrogate;
    return kSuccess;
  }
  if ((b0 & 0xF8) == 0xF0) {
    if (i + 4 > len) return kTooShort;
    const uint32_t b1 = at(i + 1), b2 = at(i + 2), b3 = at(i + 3);
    if ((b1 & 0xC0) != 0x80 || (b2 & 0xC0) != 0x80 || (b3 & 0xC0) != 0x80) return kTooShort;
    const uint32_t cp = ((b0 & 0x07) << 18) | ((b1 & 0x3F) << 12) | ((b2 & 0x3F) << 6) | (b3 & 0x3F);
    if (cp <= 0xFFFF) return kOverlong;
    if (cp > 0x10FFFF) return kTooLarge;
    return kSuccess;
  }
  return kHeaderBits;
}

// ---------------------------------------------------------------------------------------------
// UTF-16LE.  A granule is 8 units u[0..7]; `pu` / `nu` are the units before / after it.
// ---------------------------------------------------------------------------------------------
B200_HD uint32_t u16_unit(const uint32_t w[4], int i) { return (w[i >> 1] >> (16 * (i & 1))) & 0xFFFFu; }

// Bytes each unit contributes: 1 + [u > 0x7F] + [0x800 <= u, not a surrogate]; a surrogate counts 2
// (reference src/scalar/utf16.h:80-94).
B200_HD uint32_t u16_utf8_bytes(uint32_t u) {
  return 1u + (u > 0x7Fu) + (uint32_t)((u > 0x7FFu && u <= 0xD7FFu) || u >= 0xE000u);
}
// Local surrogate verdict at unit i (reference src/scalar/utf16_to_utf8/utf16_to_utf8.h:126-141,
// src/scalar/utf16.h:44-60): a low surrogate must follow a high one; a high one must be followed by
// a low one inside the buffer.  has_prev / has_next say whether pu / nu exist.
B200_HD bool u16_bad(uint32_t u, uint32_t pu, bool has_prev, uint32_t nu, bool has_next) {
  if ((u & 0xFC00u) == 0xDC00u) return !(has_prev && (pu & 0xFC00u) == 0xD800u);
  if ((u & 0xFC00u) == 0xD800u) return !(has_next && (nu & 0xFC00u) == 0xDC00u);
  return false;
}
// ---------------------------------------------------------------------------------------------
// Exact pairing screen for 8 UTF-16 units (4 little-endian words; BE callers swap first): zero iff the low
// surrogates are exactly the units behind the high surrogates, the unit before the group (upper half of pw) and
// the unit after it (lower half of nw) included.  Zero implies u16_bad() is false for all 8 units; non-zero means
// one of them is bad OR the unit before is a lone high / the unit after a lone low surrogate (reported by their
// own groups): callers treat non-zero as "search this group unit by unit".
// ---------------------------------------------------------------------------------------------
// Two tag tests per word (high / low surrogate).  Same answer as u16_pairing_screen below, which tests each word once;
// to_well_formed_utf16 keeps this form: with the other one it ran at 2470 instead of 2990 GB/s (B200, measured twice).
B200_HD uint32_t u16_is_tag(uint32_t x, uint32_t tag) {  // 0x8000 per 16-bit half with (half & 0xFC00) == tag's half
  const uint32_t z = (x & 0xFC00FC00u) ^ tag;
  return ~(((z & 0x7FFF7FFFu) + 0x7FFF7FFFu) | z) & 0x80008000u;
}
B200_HD uint32_t u16_pairing_screen_tags(const uint32_t w[4], uint32_t pw, uint32_t nw) {
  uint32_t wrong = 0, hprev = u16_is_tag(pw, 0xD800D800u);
  for (int k = 0; k < 4; k++) {
    const uint32_t H = u16_is_tag(w[k], 0xD800D800u), L = u16_is_tag(w[k], 0xDC00DC00u);
    wrong |= ((H << 16) | (hprev >> 16)) ^ L;  // the high-surrogate flags moved to the unit behind them
    hprev = H;
  }
  return wrong | ((hprev >> 16) ^ (u16_is_tag(nw, 0xDC00DC00u) & 0x8000u));
}

B200_HD uint32_t u16_sur_flags(uint32_t x) {  // 0x8000 per 16-bit half in D800..DFFF
  const uint32_t z = (x & 0xF800F800u) ^ 0xD800D800u;
  return ~(((z & 0x7FFF7FFFu) + 0x7FFF7FFFu) | z) & 0x80008000u;
}
// One surrogate test per word; bit 10 of a surrogate tells low from high (x << 5 moves it onto the flag's bit).
B200_HD uint32_t u16_pairing_screen(const uint32_t w[4], uint32_t pw, uint32_t nw) {
  uint32_t wrong = 0, hprev = u16_sur_flags(pw) & ~(pw << 5);
  for (int k = 0; k < 4; k++) {
    const uint32_t S = u16_sur_flags(w[k]), T = w[k] << 5;
    const uint32_t H = S & ~T, L = S & T;
    wrong |= ((H << 16) | (hprev >> 16)) ^ L;  // the high-surrogate flags moved to the unit behind them
    hprev = H;
  }
  return wrong | ((hprev >> 16) ^ (u16_sur_flags(nw) & (nw << 5) & 0x8000u));
}

// ---------------------------------------------------------------------------------------------
// UTF-8 -> Latin-1 (reference src/scalar/utf8_to_latin1/utf8_to_latin1.h:83-149) on 64 bytes = 16 words at once:
// returns the number of continuation bytes (64 minus it is the output length) and sets *suspect unless every byte
// >= 0xC0 is C2 / C3 and the continuation bytes are exactly the bytes behind those leads (pb / nb: the byte before
// and after the 64).  *suspect == false  =>  the reference's walk finds no error in these 64 bytes.
// ---------------------------------------------------------------------------------------------
B200_HD uint32_t u8l1_screen64(const uint32_t w[16], uint32_t pb, uint32_t nb, bool *suspect) {
  uint32_t acc = 0, wrong = 0;
  uint32_t lead_prev = pb >= 0xC0u ? 0x80000000u : 0u;
  for (int j = 0; j < 16; j++) {
    const uint32_t hi = w[j] & 0x80808080u, b6 = (w[j] << 1) & 0x80808080u;
    const uint32_t cont = hi & ~b6, lead = hi & b6;
    const uint32_t x = (w[j] & 0xFEFEFEFEu) ^ 0xC2C2C2C2u;     // zero bytes <=> C2 / C3
    const uint32_t nz = ((x & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | x;  // bit 7 set <=> byte of x non-zero
    wrong |= lead & nz;
    wrong |= ((lead << 8) | (lead_prev >> 24)) ^ cont;          // the lead flags moved to the byte behind them
    lead_prev = lead;
    acc += cont >> 7;                                           // packed byte counters, <= 16 each
  }
  wrong |= (lead_prev >> 24) ^ ((nb & 0xC0u) == 0x80u ? 0x80u : 0u);  // the last lead needs nb to continue it
  *suspect = wrong != 0u;
  return (acc * 0x01010101u) >> 24;
}
// Latin-1 -> UTF-8: bytes >= 0x80 among 64 (each becomes two bytes)
B200_HD uint32_t l1_high_count64(const uint32_t w[16]) {
  uint32_t acc = 0;
  for (int j = 0; j < 16; j++) acc += (w[j] >> 7) & 0x01010101u;
  return (acc * 0x01010101u) >> 24;
}

// Four bytes of VALID Latin-1-range UTF-8 at once (the lane passed u8l1_screen64): byte k of the result is the Latin-1
// byte of the character that STARTS at input byte k (itself if ASCII, else two bits of the C2/C3 lead and six of the
// byte behind it — wn supplies the byte behind byte 3); *keep7 has bit 7 of byte k set unless input byte k is a
// continuation byte (those produce nothing).
B200_HD uint32_t u8l1_word(uint32_t w, uint32_t wn, uint32_t *keep7) {
  const uint32_t nxt = (w >> 8) | (wn << 24);
  const uint32_t b6 = (w << 1) & 0x80808080u;
  const uint32_t m = ((w & b6) >> 7) * 0xFFu;  // 0xFF per byte >= 0xC0
  const uint32_t val = ((w & 0x03030303u) << 6) | (nxt & 0x3F3F3F3Fu);
  *keep7 = ((w & ~b6) & 0x80808080u) ^ 0x80808080u;
  return (w & ~m) | (val & m);
}
// Four Latin-1 bytes at once: *first = the first UTF-8 byte of each (itself, or C2/C3), *second = the continuation
// byte of each (meaningful where the byte is >= 0x80); returns 1 per byte >= 0x80 at bits 0, 8, 16, 24.
B200_HD uint32_t l1u8_word(uint32_t w, uint32_t *first, uint32_t *second) {
  const uint32_t hi = (w >> 7) & 0x01010101u, m = hi * 0xFFu;
  *first = (w & ~m) | ((((w >> 6) & 0x03030303u) | 0xC0C0C0C0u) & m);
  *second = (w & 0x3F3F3F3Fu) | 0x80808080u;
  return hi;
}

// ---------------------------------------------------------------------------------------------
// Base64 character classes: 0..63 sextet, 64 ASCII whitespace (' ' \t \n \r \f), 255 anything else.
// Equals the three 256-entry tables at reference src/tables/base64_tables.h:791-849
// (to_base64_value / to_base64_url_value / to_base64_default_or_url_value).
// ---------------------------------------------------------------------------------------------
B200_HD uint32_t b64_class(uint32_t c, bool url, bool both) {
  if (c - 'A' < 26u) return c - 'A';
  if (c - 'a' < 26u) return c - 'a' + 26u;
  if (c - '0' < 10u) return c - '0' + 52u;
  if (c == ' ' || c == '\t' || c == '\n' || c == '\r' || c == '\f') return 64u;
  if (c == '+') return (both || !url) ? 62u : 255u;
  if (c == '/') return (both || !url) ? 63u : 255u;
  if (c == '-') return (both || url) ? 62u : 255u;
  if (c == '_') return (both || url) ? 63u : 255u;
  return 255u;
}

// Four sextets (one per byte of w, the first in the lowest byte) -> the three bytes of their quantum in the low 24 bits,
// first byte lowest (reference src/scalar/base64.h:100-131: triple = s0 << 18 | s1 << 12 | s2 << 6 | s3, big-endian).
B200_HD uint32_t b64_pack_quantum(uint32_t w) {
  const uint32_t t1 = (w & 0x00FF00FFu) * 64u + ((w >> 8) & 0x00FF00FFu);  // s0:s1 | s2:s3, 12 bits each
  const uint32_t x = (t1 & 0xFFFFu) * 4096u + (t1 >> 16);                 // 24 bits, first byte on top
  return prmt(x, 0u, 0x4012);                                              // first byte lowest
}
// Sixteen sextets (four words) -> twelve bytes (three words), stream order.
B200_HD void b64_pack_quanta4(const uint32_t w[4], uint32_t out[3]) {
  const uint32_t y0 = b64_pack_quantum(w[0]), y1 = b64_pack_quantum(w[1]), y2 = b64_pack_quantum(w[2]), y3 = b64_pack_quantum(w[3]);
  out[0] = prmt(y0, y1, 0x4210);
  out[1] = prmt(y1, y2, 0x5421);
  out[2] = prmt(y2, y3, 0x6542);
}

// Output bytes fully determined by the first R sextets of the stream: floor(6R/8) — i.e. 3 per complete
// quantum, +1 for 2 leftover sextets, +2 for 3 (reference src/scalar/base64.h:160-200, loose mode).
B200_HD uint64_t b64_bytes_from_sextets(uint64_t r) { return (r * 6) >> 3; }

// Last-chunk / padding rules applied after the whole stream was scanned without meeting an invalid
// character (reference src/scalar/base64.h:138-200 for the partial chunk, src/generic/base64.h:228-244 /
// src/fallback/implementation.cpp:611-619 for the padding consistency check, and
// src/generic/base64.h:74-85 for input that is empty once trailing whitespace / '=' are stripped).
//   srclen        input length after stripping trailing whitespace and up to two '=' (loose/strict modes)
//   equalsigns    number of '=' stripped (0..2), equallocation = index of the first of them
//   V             number of sextet characters in [0, srclen)
//   tail_val[k]   value of the (k+1)-th sextet character counted from the end, tail_pos[k] its index
//                 (only the first V%4 entries are read)
// Output bytes for (error == SUCCESS / BASE64_INPUT_REMAINDER) are a prefix of what the decode kernel has
// already written: floor(6V/8) bytes (b64_bytes_from_sextets).
B200_HD void b64_finish(uint64_t srclen, uint32_t equalsigns, uint64_t equallocation, uint64_t V, bool garbage,
                        uint64_t last_chunk, const uint32_t tail_val[3], const uint64_t tail_pos[3], int *error,
                        uint64_t *in_count, uint64_t *out_count) {
  const uint64_t kStrict = 1, kStopBeforePartial = 2;
  if (srclen == 0) {
    *in_count = 0;
    *out_count = 0;
    *error = kSuccess;
    if (!garbage && equalsigns > 0) {
      if (last_chunk == kStrict) *error = kBase64InputRemainder;
      else if (last_chunk == kStopBeforePartial) *error = kSuccess;
      else { *error = kInvalidBase64Character; *in_count = equallocation; }
    }
    return;
  }
  const uint32_t idx = (uint32_t)(V & 3);
  const uint64_t full = 3 * (V >> 2);
  *in_count = srclen;
  *out_count = full;
  *error = kSuccess;
  if (!garbage && last_chunk == kStrict && idx != 1 && ((idx + equalsigns) & 3) != 0) {
    *error = kBase64InputRemainder;
    return;
  }
  if (!garbage && last_chunk == kStopBeforePartial && ((idx + equalsigns) & 3) != 0) {
    *in_count = idx ? tail_pos[idx - 1] : srclen;  // first character of the partial chunk
    return;
  }
  if (idx == 2) {
    if (!garbage && last_chunk == kStrict && (tail_val[0] & 0x0F)) { *error = kBase64ExtraBits; return; }
    *out_count = full + 1;
  } else if (idx == 3) {
    if (!garbage && last_chunk == kStrict && (tail_val[0] & 0x03)) { *error = kBase64ExtraBits; return; }
    *out_count = full + 2;
  } else if (idx == 1 && !garbage && last_chunk != kStopBeforePartial) {
    *error = kBase64InputRemainder;
    return;
  }
  if (last_chunk != kStopBeforePartial && equalsigns > 0 && !garbage) {
    if ((*out_count % 3 == 0) || ((*out_count % 3) + 1 + equalsigns != 4)) {
      *error = kInvalidBase64Character;
      *in_count = equallocation;
    }
  }
}

}  // namespace b200
