// swar.h — per-granule (16-byte) logic shared by the sm_100a kernels.
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
B200_HD uint32_t back1(uint32_t prev, uint32_t cur) { return prmt(prev, cur, 0x6543); }
B200_HD uint32_t back2(uint32_t prev, uint32_t cur) { return prmt(prev, cur, 0x5432); }
B200_HD uint32_t back3(uint32_t prev, uint32_t cur) { return prmt(prev, cur, 0x4321); }
B200_HD uint32_t fwd1(uint32_t cur, uint32_t next) { return prmt(cur, next, 0x4321); }
B200_HD uint32_t fwd2(uint32_t cur, uint32_t next) { return prmt(cur, next, 0x5432); }
B200_HD uint32_t fwd3(uint32_t cur, uint32_t next) { return prmt(cur, next, 0x6543); }

// Collapse a bit-7 mask word (bits 7,15,23,31) into 4 contiguous bits (bit k = byte k).
B200_HD uint32_t mask4(uint32_t m) { return ((m >> 7) * 0x00204081u) >> 21 & 0xF; }
// Expand 4 bits into a bit-7 mask word.
B200_HD uint32_t unmask4(uint32_t b) {
  return ((b & 1) << 7) | ((b & 2) << 14) | ((b & 4) << 21) | ((b & 8) << 28);
}

// ---------------------------------------------------------------------------------------------
// UTF-8 byte classes
// ---------------------------------------------------------------------------------------------
struct U8Class {
  uint32_t cont;  // 10xxxxxx
  uint32_t l2;    // >= 0xC0 (any lead or invalid header)
  uint32_t l3;    // >= 0xE0
  uint32_t l4;    // >= 0xF0
};

B200_HD U8Class u8_classify(uint32_t w) {
  const uint32_t s1 = w << 1, s2 = w << 2, s3 = w << 3;  // bit (7-k) of every byte moved to bit 7
  U8Class c;
  c.l2 = w & s1 & kH;
  c.l3 = c.l2 & s2;
  c.l4 = c.l3 & s3;
  c.cont = w & ~s1 & kH;
  return c;
}

// Non-continuation bytes (code-point starts), reference src/scalar/utf8.h:230-241: (int8_t)b > -65.
B200_HD uint32_t u8_noncont(uint32_t w) { return (~w | (w << 1)) & kH; }
// Bytes >= 0xF0, reference src/scalar/utf8.h:243-255 (the second unit of a surrogate pair).
B200_HD uint32_t u8_ge_f0(uint32_t w) { return w & (w << 1) & (w << 2) & (w << 3) & kH; }

// The bytes whose *second* byte range is restricted, packed one per bit:
//   bit7 = 0xE0 (next must be A0..BF), bit6 = 0xED (next must be 80..9F),
//   bit5 = 0xF0 (next must be 90..BF), bit4 = 0xF4 (next must be 80..8F);
// `bad` = bytes that can never appear: 0xC0, 0xC1, 0xF5..0xFF.
// (Rules: reference src/scalar/utf8.h:133-185 — OVERLONG / SURROGATE / TOO_LARGE range checks.)
B200_HD uint32_t u8_special(uint32_t w, const U8Class &c, uint32_t *bad) {
  const uint32_t x = w & 0x7F7F7F7Fu;             // 7-bit payload; x + k never carries across bytes (k <= 0x3E)
  const uint32_t geE1 = x + 0x1F1F1F1Fu;          // bit7 set iff byte >= 0xE1 (given bit7 of w)
  const uint32_t geED = x + 0x13131313u;
  const uint32_t geEE = x + 0x12121212u;
  const uint32_t geF1 = x + 0x0F0F0F0Fu;
  const uint32_t geF4 = x + 0x0C0C0C0Cu;
  const uint32_t geF5 = x + 0x0B0B0B0Bu;
  const uint32_t geC2 = x + 0x3E3E3E3Eu;
  const uint32_t isE0 = c.l3 & ~geE1;
  const uint32_t isED = c.l3 & geED & ~geEE;
  const uint32_t isF0 = c.l4 & ~geF1;
  const uint32_t isF4 = c.l4 & geF4 & ~geF5;
  *bad = (c.l2 & ~geC2) | (c.l4 & geF5);
  return isE0 | (isED >> 1) | (isF0 >> 2) | (isF4 >> 3);
}

// Detects (does not classify) UTF-8 errors in one granule.  Returns non-zero iff some byte of the
// granule violates the structure ("must be continuation" != "is continuation", looking back 3 bytes
// into `prev`) or a second-byte range rule.  A sequence truncated by the end of the buffer shows up
// as a violation on the zero filler byte that follows it — or, if no granule follows, is caught by
// u8_incomplete_tail().  A non-zero return only says "run u8_verdict on bytes [lo-3, hi)".
B200_HD uint32_t u8_check_granule(const uint32_t w[4], uint32_t prev) {
  U8Class pc = u8_classify(prev);
  uint32_t pbad;
  uint32_t pk = u8_special(prev, pc, &pbad);
  uint32_t err = 0;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int k = 0; k < 4; k++) {
    const uint32_t cur = w[k];
    const U8Class c = u8_classify(cur);
    const uint32_t must = back1(pc.l2, c.l2) | back2(pc.l3, c.l3) | back3(pc.l4, c.l4);
    uint32_t bad;
    const uint32_t kk = u8_special(cur, c, &bad);
    const uint32_t k1 = back1(pk, kk);           // special-lead flags of the previous byte
    const uint32_t c5 = cur << 2, c4 = cur << 3; // bit5 / bit4 of the current byte at bit 7
    const uint32_t c54 = c5 | c4;
    const uint32_t f1 = (k1 & ~c5) | ((k1 << 1) & c5);            // E0 then 80..9F | ED then A0..BF
    const uint32_t f2 = ((k1 << 2) & ~c54) | ((k1 << 3) & c54);   // F0 then 80..8F | F4 then 90..BF
    err |= (must ^ c.cont) | bad | ((f1 | f2) & kH);
    pc = c;
    pk = kk;
  }
  return err;
}

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
// UTF-8 -> UTF-16LE / UTF-32 emission, one output element per *position*:
//   UTF-16: every non-continuation byte emits one unit (the high surrogate for a 4-byte lead), and
//           the byte right after a byte >= 0xF0 emits the low surrogate.  Summed over the buffer this
//           is exactly utf16_length_from_utf8 (reference src/scalar/utf8.h:243-255) for valid input
//           and never more for invalid input, so an output buffer sized by that query is never overrun
//           (reference tests/convert_utf8_to_utf16le_tests.cpp:23-52).
//   UTF-32: every non-continuation byte emits one word (count_utf8, src/scalar/utf8.h:230-241).
// Values follow reference src/scalar/utf8_to_utf16/utf8_to_utf16.h:154-242 and
// src/scalar/utf8_to_utf32/utf8_to_utf32.h:128-200.
// ---------------------------------------------------------------------------------------------
B200_HD void u8_emit16_masks(const uint32_t w[4], uint32_t prev, uint32_t em[4]) {
  uint32_t pf = u8_ge_f0(prev);
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int k = 0; k < 4; k++) {
    const uint32_t f = u8_ge_f0(w[k]);
    em[k] = u8_noncont(w[k]) | back1(pf, f);
    pf = f;
  }
}
B200_HD void u8_emit32_masks(const uint32_t w[4], uint32_t em[4]) {
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int k = 0; k < 4; k++) em[k] = u8_noncont(w[k]);
}

// Byte p (-4 <= p < 20) of the 24-byte window {prev, w[0..3], next}.
#define B200_WBYTE(p) ((win[((p) + 4) >> 2] >> (8 * (((p) + 4) & 3))) & 0xFFu)

template <class Sink>
B200_HD void u8_emit16_granule(const uint32_t w[4], uint32_t prev, uint32_t next, const uint32_t em[4], Sink &&sink) {
  const uint32_t win[6] = {prev, w[0], w[1], w[2], w[3], next};
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int p = 0; p < 16; p++) {
    if (!((em[p >> 2] >> (8 * (p & 3) + 7)) & 1)) continue;
    const uint32_t b0 = B200_WBYTE(p);
    uint32_t unit;
    if (b0 < 0x80) {
      unit = b0;
    } else if (b0 < 0xC0) {  // second byte of a 4-byte sequence: low surrogate from bytes 3 and 4
      unit = 0xDC00u | ((B200_WBYTE(p + 1) & 0x0F) << 6) | (B200_WBYTE(p + 2) & 0x3F);
    } else if (b0 < 0xE0) {
      unit = ((b0 & 0x1F) << 6) | (B200_WBYTE(p + 1) & 0x3F);
    } else if (b0 < 0xF0) {
      unit = ((b0 & 0x0F) << 12) | ((B200_WBYTE(p + 1) & 0x3F) << 6) | (B200_WBYTE(p + 2) & 0x3F);
    } else {  // high surrogate: (cp >> 10) - 0x40
      const uint32_t v = ((b0 & 0x07) << 8) | ((B200_WBYTE(p + 1) & 0x3F) << 2) | ((B200_WBYTE(p + 2) >> 4) & 0x03);
      unit = 0xD800u + ((v - 0x40u) & 0x3FFu);
    }
    sink((uint16_t)unit);
  }
}

template <class Sink>
B200_HD void u8_emit32_granule(const uint32_t w[4], uint32_t next, const uint32_t em[4], Sink &&sink) {
  const uint32_t win[6] = {0u, w[0], w[1], w[2], w[3], next};
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int p = 0; p < 16; p++) {
    if (!((em[p >> 2] >> (8 * (p & 3) + 7)) & 1)) continue;
    const uint32_t b0 = B200_WBYTE(p);
    uint32_t cp;
    if (b0 < 0x80) {
      cp = b0;
    } else if (b0 < 0xE0) {
      cp = ((b0 & 0x1F) << 6) | (B200_WBYTE(p + 1) & 0x3F);
    } else if (b0 < 0xF0) {
      cp = ((b0 & 0x0F) << 12) | ((B200_WBYTE(p + 1) & 0x3F) << 6) | (B200_WBYTE(p + 2) & 0x3F);
    } else {
      cp = ((b0 & 0x07) << 18) | ((B200_WBYTE(p + 1) & 0x3F) << 12) | ((B200_WBYTE(p + 2) & 0x3F) << 6) |
           (B200_WBYTE(p + 3) & 0x3F);
    }
    sink(cp);
  }
}

// ---------------------------------------------------------------------------------------------
// Branch-free SWAR transcoder UTF-8 -> UTF-16LE, four byte positions per 32-bit word, fused with the
// validation detector.  Every position computes the unit it WOULD emit from its own byte and the payloads
// of the next two bytes (x1, x2):
//     ASCII         b                                   2-byte lead   (b&1F)<<6 | x1
//     3-byte lead   (b&0F)<<12 | x1<<6 | x2             4-byte lead   0xD800 | ((cp>>16)-1)<<6 | (x1&0F)<<2 | x2>>4
//     byte after a 4-byte lead (its own x1,x2 are bytes 3,4 of the sequence)   0xDC00 | (x1&0F)<<6 | x2
// as a low-byte plane and a high-byte plane, then interleaves the planes into units.  Which positions
// really emit is the caller's mask (non-continuation bytes + the byte after a byte >= 0xF0), so the number
// of units per byte equals utf16_length_from_utf8's count (reference src/scalar/utf8.h:243-255).
// Values: reference src/scalar/utf8_to_utf16/utf8_to_utf16.h:154-242.
// The detector flags exactly the invalid inputs (structure: "must be continuation" vs "is continuation";
// ranges: on the decoded bits — overlong 2/3/4-byte, surrogates, > U+10FFFF, 0xF8..0xFF), see
// u8_check_granule for how a flag is turned into the exact (error, position).
// ---------------------------------------------------------------------------------------------
struct U8Carry {       // class masks of the previous word (what the current word looks back at)
  uint32_t l2, l3, l4;
};
struct U8Word16 {
  uint32_t u01, u23;   // candidate units of positions 0,1 and 2,3 (low half = lower position)
  uint32_t emit;       // bit-7 mask: positions that emit a unit
  uint32_t err;        // bit-7 mask: validation detector (non-zero => run the exact locator)
};

B200_HD U8Carry u8_carry_of(uint32_t w) {
  const U8Class c = u8_classify(w);
  U8Carry k;
  k.l2 = c.l2; k.l3 = c.l3; k.l4 = c.l4;
  return k;
}

// Full-byte (0xFF / 0x00) mask from a bit-7 mask: PRMT with the sign-replicate selector.
B200_HD uint32_t fullmask(uint32_t m) {
#if defined(__CUDA_ARCH__)
  uint32_t r;  // prmt's selector msb = "replicate the sign of the selected byte" (__byte_perm ignores that bit)
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(m), "r"(0u), "r"(0xBA98u));
  return r;
#else
  return ((m >> 7) & 0x01010101u) * 0xFFu;
#endif
}
// bit 7 of each byte set iff (t & 0xFE) != 0, for t with bit 0 clear in every byte... generalised below:
// nz7(t): t must have bit 0 of every byte clear; returns bit-7 mask of the non-zero bytes.
B200_HD uint32_t nz7(uint32_t t) { return ((t >> 1) + 0x7F7F7F7Fu) & kH; }

template <bool VALIDATE>
B200_HD U8Word16 u8_to_utf16_word(uint32_t w, uint32_t xnext, U8Carry &carry) {
  const uint32_t s1 = w << 1, s2 = w << 2, s3 = w << 3;
  const uint32_t l2 = w & s1 & kH, l3 = l2 & s2, l4 = l3 & s3;
  const uint32_t cont = w & ~s1 & kH;
  const uint32_t x = w & 0x3F3F3F3Fu;
  const uint32_t x1 = prmt(x, xnext, 0x4321), x2 = prmt(x, xnext, 0x5432);
  const uint32_t afterf0 = back1(carry.l4, l4);  // the byte after a byte >= 0xF0 carries the low surrogate
  // ---- low / high byte planes per class ----
  const uint32_t lo2 = ((w << 6) & 0xC0C0C0C0u) | x1;
  const uint32_t hi2 = (w >> 2) & 0x07070707u;
  const uint32_t x1r2 = x1 >> 2;
  const uint32_t lo3 = ((x1 << 6) & 0xC0C0C0C0u) | x2;                   // also the low surrogate's low byte
  const uint32_t hi3 = (s2 << 2 & 0xF0F0F0F0u) | (x1r2 & 0x0F0F0F0Fu);   // (w<<4 & F0) | (x1>>2 & 0F)
  const uint32_t v = (s2 & 0x1C1C1C1Cu) | ((x1 >> 4) & 0x03030303u);     // cp >> 16 (5 bits) at 4-byte leads
  const uint32_t v1 = v + 0x7F7F7F7Fu;                                   // low 7 bits: v - 1 ; bit 7: v >= 1
  const uint32_t lo4 = ((x1 << 2) & 0x3C3C3C3Cu) | ((((v1 << 6) & 0xC0C0C0C0u) | ((x2 >> 4) & 0x3F3F3F3Fu)) & 0xC3C3C3C3u);
  const uint32_t l4F = fullmask(l4), contF = fullmask(cont), hiF = fullmask(w), is2F = fullmask(l2 & ~l3);
  // surrogate high bytes: 0xD8 | ((cp>>16)-1)>>2 at the lead, 0xDC | (x1>>2 & 3) at the byte after it
  const uint32_t sp = (((v1 >> 2) & l4F) | (x1r2 & ~l4F)) & 0x03030303u;
  const uint32_t hs = (sp | 0xD8D8D8D8u) | (contF & 0x04040404u);
  const uint32_t surrF = l4F | contF;
  const uint32_t lo = (w & ~hiF) | (hiF & ((is2F & lo2) | (~is2F & ((l4F & lo4) | (~l4F & lo3)))));
  const uint32_t hi = hiF & ((surrF & hs) | (~surrF & ((is2F & hi2) | (~is2F & hi3))));
  U8Word16 r;
  r.u01 = prmt(lo, hi, 0x5140);
  r.u23 = prmt(lo, hi, 0x7362);
  r.emit = (~cont & kH) | afterf0;
  r.err = 0;
  if (VALIDATE) {
    const uint32_t must = back1(carry.l2, l2) | back2(carry.l3, l3) | back3(carry.l4, l4);
    const uint32_t is2 = l2 & ~l3, is3 = l3 & ~l4;
    const uint32_t over2 = is2 & ~(((w & 0x1E1E1E1Eu) + 0x7F7F7F7Fu));           // C0 / C1
    const uint32_t h8 = hi3 & 0xF8F8F8F8u;
    const uint32_t over3 = is3 & ~nz7(h8);                                         // E0 80..9F
    const uint32_t surr3 = is3 & ~nz7(h8 ^ 0xD8D8D8D8u);                           // ED A0..BF
    const uint32_t bad4 = l4 & (nz7((v1 ^ kH) & 0xF0F0F0F0u) | (s3 << 1));         // cp>>16 not in 1..16, or F8..FF
    r.err = (must ^ cont) | over2 | over3 | surr3 | bad4;
  }
  carry.l2 = l2; carry.l3 = l3; carry.l4 = l4;
  return r;
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
// UTF-8 bytes of unit u (reference src/scalar/utf16_to_utf8/utf16_to_utf8.h:105-148).  A high surrogate
// emits the first two bytes of the 4-byte sequence, the low surrogate the last two (it needs the two low
// bits of (cp >> 10), taken from the preceding unit).
template <class Sink>
B200_HD void u16_emit8_unit(uint32_t u, uint32_t pu, Sink &&sink) {
  if (u < 0x80u) {
    sink((uint8_t)u);
  } else if (u < 0x800u) {
    sink((uint8_t)(0xC0u | (u >> 6)));
    sink((uint8_t)(0x80u | (u & 0x3Fu)));
  } else if ((u & 0xF800u) != 0xD800u) {
    sink((uint8_t)(0xE0u | (u >> 12)));
    sink((uint8_t)(0x80u | ((u >> 6) & 0x3Fu)));
    sink((uint8_t)(0x80u | (u & 0x3Fu)));
  } else if ((u & 0xFC00u) == 0xD800u) {
    const uint32_t t = (u & 0x3FFu) + 0x40u;  // cp >> 10
    sink((uint8_t)(0xF0u | (t >> 8)));
    sink((uint8_t)(0x80u | ((t >> 2) & 0x3Fu)));
  } else {
    const uint32_t t = (pu & 0x3FFu) + 0x40u;
    sink((uint8_t)(0x80u | ((t & 3u) << 4) | ((u >> 6) & 0x0Fu)));
    sink((uint8_t)(0x80u | (u & 0x3Fu)));
  }
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
