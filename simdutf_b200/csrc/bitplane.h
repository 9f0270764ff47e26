// bitplane.h — bit-plane ("transposed") formulation of the UTF-8 hot path for sm_100a.
//
// A *block* is 32 consecutive input bytes held by ONE thread as eight little-endian 32-bit words.  The block is
// transposed into eight bit planes B[0..7]: bit p of plane k is bit k of byte p.  In that form every bitwise
// instruction classifies / validates / assembles 32 byte positions at once, where the byte-SWAR form of
// swar.h handles 4.  "The byte before" is a 1-bit funnel shift of a plane (the bits shifted in come from
// the previous block's plane).  The 16 planes of the candidate UTF-16 unit of every position are transposed
// back into sixteen words of two units each.
//
// Like swar.h everything is `__host__ __device__` so tests/host/ runs exactly this code on the CPU against
// the oracle; nothing in the product path calls it on the host.
//
// Semantics (reference file:line):
//   validation rules         src/scalar/utf8.h:102-200 (detected here, located exactly by swar.h:u8_verdict)
//   UTF-16 unit values       src/scalar/utf8_to_utf16/utf8_to_utf16.h:154-242
//   UTF-32 values            src/scalar/utf8_to_utf32/utf8_to_utf32.h:128-200
//
// Emission rule ("end of character"): position p emits the unit of the character that ENDS at p, i.e. when
// byte p+1 is not a continuation byte (the zero filler past the end of the buffer counts as one); a 4-byte
// character additionally emits its high surrogate at its THIRD byte (the position two after a byte >= 0xF0),
// where lead, second and third byte are all behind or at the position.  All data dependencies therefore
// look backwards (<= 3 bytes) except the one "is the next byte a continuation" bit.
// For a buffer whose first byte is not a continuation byte the number of emitting positions is
//   #non-continuation bytes in (0, len]  +  #bytes >= 0xF0 in [0, len-2)  <=  utf16_length_from_utf8(buf),
// with equality for valid input (reference src/scalar/utf8.h:243-255), so an output buffer sized by that query
// is never overrun.  A buffer that starts with a continuation byte is invalid at position 0 (TOO_LONG); the
// kernels emit nothing for it.
#pragma once
#include "swar.h"

namespace b200 {
namespace bp {

// (hi:lo) << n, upper word: the plane `hi` moved n positions later, filled from the top of `lo`.
B200_HD uint32_t fsl(uint32_t lo, uint32_t hi, int n) {
#if defined(__CUDA_ARCH__)
  return __funnelshift_l(lo, hi, n);
#else
  return (hi << n) | (lo >> (32 - n));
#endif
}
// (hi:lo) >> n, lower word.
B200_HD uint32_t fsr(uint32_t lo, uint32_t hi, int n) {
#if defined(__CUDA_ARCH__)
  return __funnelshift_r(lo, hi, n);
#else
  return (lo >> n) | (hi << (32 - n));
#endif
}

// (x & m) | (y & ~m) as ONE three-input logic op: written with two different immediates the compiler does not
// see that the masks are complements and spends three.
B200_HD uint32_t bitsel(uint32_t x, uint32_t y, uint32_t m) {
#if defined(__CUDA_ARCH__)
  uint32_t r;
  asm("lop3.b32 %0, %1, %2, %3, 0xE4;" : "=r"(r) : "r"(x), "r"(y), "r"(m));
  return r;
#else
  return (x & m) | (y & ~m);
#endif
}

// x >> S through the multiplier: on sm_100 LOP3/SHF/PRMT share the half-rate ALU pipe, which bounds these kernels,
// while IMAD / IMAD.HI issue on the FMA pipe (tools/ubench/pipes.cu: LOP3 + IMAD interleaved run at 3.7 warp
// instructions per clock per SM, either alone at 2.0; IMAD.HI at 1.0).
template <int S>
B200_HD uint32_t shr_fma(uint32_t x) {
#if defined(__CUDA_ARCH__)
  return __umulhi(x, 1u << (32 - S));
#else
  return x >> S;
#endif
}

// Exchange one word-index bit with one bit-index bit: a holds the elements whose word-index bit is 0.
//   a' = a's low-group bits in place, b's low-group bits moved up by s
//   b' = a's high-group bits moved down by s, b's high-group bits in place
template <int S, uint32_t M0>
B200_HD void dswap(uint32_t &a, uint32_t &b) {
  const uint32_t a2 = bitsel(a, b << S, M0);
  const uint32_t b2 = bitsel(shr_fma<S>(a), b, M0);
  a = a2;
  b = b2;
}
B200_HD void dswap_bytes(uint32_t &a, uint32_t &b) {  // S = 8, M0 = 0x00FF00FF as two byte permutes
  const uint32_t a2 = prmt(a, b, 0x6240);
  const uint32_t b2 = prmt(a, b, 0x7351);
  a = a2;
  b = b2;
}
// 4x4 byte transpose: afterwards a = byte 0 of (a,b,c,d), b = byte 1 of each, ...
B200_HD void tr4x4(uint32_t &a, uint32_t &b, uint32_t &c, uint32_t &d) {
  const uint32_t ab_lo = prmt(a, b, 0x5140), ab_hi = prmt(a, b, 0x7362);
  const uint32_t cd_lo = prmt(c, d, 0x5140), cd_hi = prmt(c, d, 0x7362);
  a = prmt(ab_lo, cd_lo, 0x5410);
  b = prmt(ab_lo, cd_lo, 0x7632);
  c = prmt(ab_hi, cd_hi, 0x5410);
  d = prmt(ab_hi, cd_hi, 0x7632);
}

// 32 bytes (w[i] = bytes 4i..4i+3) -> 8 planes, in place: afterwards w[k] = plane k.
//   byte transposes put positions {c, c+8, c+16, c+24} (c = 0..3) and {c+4, ...} into one word each; three
//   delta-swap stages then trade the three "bit within byte" index bits for the three remaining position bits.
B200_HD void transpose_in(uint32_t (&w)[8]) {
  uint32_t e0 = w[0], e1 = w[2], e2 = w[4], e3 = w[6];
  uint32_t o0 = w[1], o1 = w[3], o2 = w[5], o3 = w[7];
  tr4x4(e0, e1, e2, e3);  // e_c: positions c + 8j      (p2 = 0)
  tr4x4(o0, o1, o2, o3);  // o_c: positions c + 4 + 8j  (p2 = 1)
  dswap<4, 0x0F0F0F0Fu>(e0, o0);
  dswap<4, 0x0F0F0F0Fu>(e1, o1);
  dswap<4, 0x0F0F0F0Fu>(e2, o2);
  dswap<4, 0x0F0F0F0Fu>(e3, o3);
  dswap<2, 0x33333333u>(e0, e2);
  dswap<2, 0x33333333u>(e1, e3);
  dswap<2, 0x33333333u>(o0, o2);
  dswap<2, 0x33333333u>(o1, o3);
  dswap<1, 0x55555555u>(e0, e1);
  dswap<1, 0x55555555u>(e2, e3);
  dswap<1, 0x55555555u>(o0, o1);
  dswap<1, 0x55555555u>(o2, o3);
  w[0] = e0; w[1] = e1; w[2] = e2; w[3] = e3;
  w[4] = o0; w[5] = o1; w[6] = o2; w[7] = o3;
}

// 16 planes (u[k] = plane k of a 16-bit value per position) -> 16 words, in place: afterwards
// u[q] = value of position q in the low half, value of position q + 16 in the high half.
B200_HD void transpose_out16(uint32_t (&u)[16]) {
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int i = 0; i < 16; i += 2) dswap<1, 0x55555555u>(u[i], u[i + 1]);
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int i = 0; i < 16; i++)
    if (!(i & 2)) dswap<2, 0x33333333u>(u[i], u[i + 2]);
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int i = 0; i < 16; i++)
    if (!(i & 4)) dswap<4, 0x0F0F0F0Fu>(u[i], u[i + 4]);
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int i = 0; i < 8; i++) dswap_bytes(u[i], u[i + 8]);
}

// dswap whose second word is known to be zero.
template <int S, uint32_t M0>
B200_HD void dswap_z(uint32_t &a, uint32_t &b) {
  const uint32_t a2 = a & M0;
  b = shr_fma<S>(a) & M0;
  a = a2;
}
// N planes (u[0..N-1]; u[N..31] are ignored and treated as zero, 16 < N <= 24) of an N-bit value per position
// -> 32 words, in place: afterwards u[q] = value of position q.  Five exchange stages; pairs of known-zero words
// are skipped, pairs with one known-zero word take the cheap form.
template <int N>
B200_HD void transpose_out_n(uint32_t (&u)[32]) {
  static_assert(N > 16 && N <= 24, "live planes");
  constexpr int L1 = (N + 1) / 2 * 2, L2 = (L1 + 3) / 4 * 4, L3 = (L2 + 7) / 8 * 8;
  static_assert(L3 == 24, "stage 3 assumes words 0..23 live");
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int i = 0; i < N; i += 2) {
    if (i + 1 < N) dswap<1, 0x55555555u>(u[i], u[i + 1]);
    else dswap_z<1, 0x55555555u>(u[i], u[i + 1]);
  }
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int i = 0; i < L1; i++) {
    if (i & 2) continue;
    if (i + 2 < L1) dswap<2, 0x33333333u>(u[i], u[i + 2]);
    else dswap_z<2, 0x33333333u>(u[i], u[i + 2]);
  }
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int i = 0; i < L2; i++) {
    if (i & 4) continue;
    if (i + 4 < L2) dswap<4, 0x0F0F0F0Fu>(u[i], u[i + 4]);
    else dswap_z<4, 0x0F0F0F0Fu>(u[i], u[i + 4]);
  }
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int i = 0; i < 8; i++) dswap_bytes(u[i], u[i + 8]);
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int i = 16; i < 24; i++) {  // partner (24..31) is zero
    const uint32_t a = u[i];
    u[i] = a & 0x00FF00FFu;
    u[i + 8] = prmt(a, 0u, 0x4341);
  }
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int i = 0; i < 16; i++) {  // halfword exchange
    const uint32_t a = u[i], b = u[i + 16];
    u[i] = prmt(a, b, 0x5410);
    u[i + 16] = prmt(a, b, 0x7632);
  }
}
B200_HD void transpose_out21(uint32_t (&u)[32]) { transpose_out_n<21>(u); }

// 32 UTF-16 units (w[i] = units 2i, 2i+1) -> 16 planes, in place.  The planes come out in SPLIT order: bit i of a
// plane belongs to unit 2i, bit 16+i to unit 2i+1 (i = 0..15) — the exact inverse of transpose_out16, which saves
// the 16 halfword permutes natural order would need.  split_pos() maps a unit index to its bit.
B200_HD void transpose_in16(uint32_t (&w)[16]) {
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int i = 0; i < 8; i++) dswap_bytes(w[i], w[i + 8]);
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int i = 0; i < 16; i++)
    if (!(i & 4)) dswap<4, 0x0F0F0F0Fu>(w[i], w[i + 4]);
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int i = 0; i < 16; i++)
    if (!(i & 2)) dswap<2, 0x33333333u>(w[i], w[i + 2]);
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int i = 0; i < 16; i += 2) dswap<1, 0x55555555u>(w[i], w[i + 1]);
}
#if defined(__CUDACC__)
__host__ __device__
#endif
constexpr int split_pos(int unit) { return (unit >> 1) + 16 * (unit & 1); }
// "The unit before" in split order: odd units look at the even unit of the same word, even units at the odd unit of
// the word before (the first one at the last unit of the previous block, bit 31 of `prev`).
B200_HD uint32_t split_prev1(uint32_t prev, uint32_t x) { return (x << 16) | ((x >> 15) & 0xFFFEu) | (prev >> 31); }

// The top four positions (28..31) of the planes of a virtual block whose last word is `pw`; the other bits
// are unspecified (only the top <= 3 bits are ever shifted into the block that follows).
B200_HD void planes_of_tail_word(uint32_t pw, uint32_t (&v)[8]) {
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int k = 0; k < 8; k++) v[k] = ((pw >> k) & 0x01010101u) * 0x10204080u;
}

// Class planes a block hands to the block after it.
struct Carry {
  uint32_t b[6];            // planes 0..5 (payload bits)
  uint32_t l2, l3, l4;      // bytes >= 0xC0 / >= 0xE0 / >= 0xF0
  uint32_t e0, ed, f0, f4;  // the four leads whose second byte has a restricted range
};

struct Classes {
  uint32_t cont, l2, l3, l4;
};
B200_HD Classes classify(const uint32_t (&B)[8]) {
  Classes c;
  c.cont = B[7] & ~B[6];
  c.l2 = B[7] & B[6];
  c.l3 = c.l2 & B[5];
  c.l4 = c.l3 & B[4];
  return c;
}

// Emitting positions of a block (before clipping to the buffer): ends of characters + third bytes of
// 4-byte sequences.  `next_noncont` = 1 iff the byte after the block is not a continuation byte;
// `prev_l4` = the >= 0xF0 plane of the previous block.
B200_HD uint32_t emit16_mask(const uint32_t (&B)[8], uint32_t prev_l4, uint32_t next_noncont) {
  const uint32_t nc = ~B[7] | B[6];
  const uint32_t l4 = B[7] & B[6] & B[5] & B[4];
  return fsr(nc, next_noncont, 1) | fsl(prev_l4, l4, 2);
}
// UTF-32 / count_utf8 flavour: one element per character, at its last byte.
B200_HD uint32_t emit32_mask(const uint32_t (&B)[8], uint32_t next_noncont) {
  const uint32_t nc = ~B[7] | B[6];
  return fsr(nc, next_noncont, 1);
}

// Seeds the carry of the first block of a run from the 4 bytes in front of it.
B200_HD Carry carry_from_word(uint32_t pw) {
  uint32_t v[8];
  planes_of_tail_word(pw, v);
  Carry c;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int k = 0; k < 6; k++) c.b[k] = v[k];
  c.l2 = v[7] & v[6];
  c.l3 = c.l2 & v[5];
  c.l4 = c.l3 & v[4];
  const uint32_t n321 = ~(v[3] | v[2] | v[1]);
  const uint32_t is3 = c.l3 & ~v[4];
  c.e0 = is3 & n321 & ~v[0];
  c.ed = is3 & v[3] & v[2] & ~v[1] & v[0];
  c.f0 = c.l4 & n321 & ~v[0];
  c.f4 = c.l4 & ~v[3] & v[2] & ~v[1] & ~v[0];
  return c;
}

// One block: validation detector + the 16 planes of the candidate UTF-16 unit of every position.
//   B      planes of the block (transpose_in)
//   c      in: class planes / payload planes of the previous block; out: those of this block
//   U      out: unit planes (meaningful at emitting positions only)
//   returns the detector plane: a set bit means "some rule is violated at or up to 3 bytes before this
//   position" (never set on valid input; every invalid input sets at least one bit within 3 positions after
//   its first offending byte, possibly on the zero filler behind the buffer — see swar.h:u8_incomplete_tail for
//   the one case without filler).
template <bool VALIDATE>
B200_HD uint32_t utf8_to_utf16_block(const uint32_t (&B)[8], Carry &c, uint32_t (&U)[16]) {
  const uint32_t b7 = B[7];
  const uint32_t cont = b7 & ~B[6];
  const uint32_t l2 = b7 & B[6], l3 = l2 & B[5], l4 = l3 & B[4];
  const uint32_t must2 = fsl(c.l3, l3, 2);  // second continuation of a 3-/4-byte sequence is due here
  const uint32_t m4e = fsl(c.l4, l4, 3);    // fourth byte of a 4-byte sequence (low surrogate)
  const uint32_t m4t = fsl(c.l4, l4, 2);    // third byte (high surrogate)
  uint32_t P[6], Q[4];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int k = 0; k < 6; k++) P[k] = fsl(c.b[k], B[k], 1);
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int k = 0; k < 4; k++) Q[k] = fsl(c.b[k], B[k], 2);
  // high surrogate: 0xD800 | ((cp >> 10) - 0x40), cp >> 10 = Q[2:0]:P[5:0]:B[5:4]
  const uint32_t bw = ~(P[5] | P[4]);  // borrow out of the two bits above P[3:0]
  const uint32_t t8 = Q[0] ^ bw;
  const uint32_t t9 = Q[1] ^ (bw & ~Q[0]);
  const uint32_t m4 = m4t | m4e;
  U[0] = (m4t & B[4]) | (~m4t & B[0]);
  U[1] = (m4t & B[5]) | (~m4t & B[1]);
  U[2] = (m4t & P[0]) | (~m4t & B[2]);
  U[3] = (m4t & P[1]) | (~m4t & B[3]);
  U[4] = (m4t & P[2]) | (~m4t & B[4]);
  U[5] = (m4t & P[3]) | (~m4t & B[5]);
  U[6] = (m4t & ~P[4]) | (~m4t & (B[6] | (b7 & P[0])));
  U[7] = (m4t & ~(P[5] ^ P[4])) | (~m4t & b7 & P[1]);
  U[8] = (m4t & t8) | (~m4t & b7 & P[2]);
  U[9] = (m4t & t9) | (~m4t & b7 & P[3]);
  U[10] = ~m4t & ((b7 & P[4]) | m4e);
  U[11] = m4 | (b7 & P[5]);
  U[12] = (must2 & Q[0]) | m4;
  U[13] = must2 & Q[1] & ~m4t;
  U[14] = (must2 & Q[2]) | m4;
  U[15] = (must2 & Q[3]) | m4;
  uint32_t err = 0;
  uint32_t e0 = 0, ed = 0, f0 = 0, f4 = 0;
  if (VALIDATE) {
    const uint32_t must = fsl(c.l2, l2, 1) | must2 | m4e;
    const uint32_t n321 = ~(B[3] | B[2] | B[1]);
    const uint32_t is3 = l3 & ~B[4];
    e0 = is3 & n321 & ~B[0];
    ed = is3 & B[3] & B[2] & ~B[1] & B[0];
    f0 = l4 & n321 & ~B[0];
    f4 = l4 & ~B[3] & B[2] & ~B[1] & ~B[0];
    const uint32_t over2 = l2 & ~B[5] & ~B[4] & n321;            // C0, C1
    const uint32_t big = l4 & (B[3] | (B[2] & (B[1] | B[0])));   // F5..FF
    const uint32_t pe0 = fsl(c.e0, e0, 1), ped = fsl(c.ed, ed, 1), pf0 = fsl(c.f0, f0, 1), pf4 = fsl(c.f4, f4, 1);
    const uint32_t rng = (pe0 & ~B[5]) | (ped & B[5]) | (pf0 & ~B[5] & ~B[4]) | (pf4 & (B[5] | B[4]));
    err = (must ^ cont) | over2 | big | rng;
  }
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int k = 0; k < 6; k++) c.b[k] = B[k];
  c.l2 = l2; c.l3 = l3; c.l4 = l4;
  c.e0 = e0; c.ed = ed; c.f0 = f0; c.f4 = f4;
  return err;
}

// UTF-32 flavour: the 21 planes of the code point of the character that ENDS at each position.
template <bool VALIDATE>
B200_HD uint32_t utf8_to_utf32_block(const uint32_t (&B)[8], Carry &c, uint32_t (&C)[32]) {
  const uint32_t b7 = B[7];
  const uint32_t cont = b7 & ~B[6];
  const uint32_t l2 = b7 & B[6], l3 = l2 & B[5], l4 = l3 & B[4];
  const uint32_t must2 = fsl(c.l3, l3, 2);
  const uint32_t m4e = fsl(c.l4, l4, 3);
  const uint32_t m34 = must2 | m4e;
  uint32_t P[6], Q[6], R[3];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int k = 0; k < 6; k++) {
    P[k] = fsl(c.b[k], B[k], 1);
    Q[k] = fsl(c.b[k], B[k], 2);
  }
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int k = 0; k < 3; k++) R[k] = fsl(c.b[k], B[k], 3);
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int k = 0; k < 6; k++) C[k] = B[k];
  C[6] = B[6] | (b7 & P[0]);
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int k = 1; k < 6; k++) C[6 + k] = b7 & P[k];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int k = 0; k < 4; k++) C[12 + k] = m34 & Q[k];
  C[16] = m4e & Q[4];
  C[17] = m4e & Q[5];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int k = 0; k < 3; k++) C[18 + k] = m4e & R[k];
  uint32_t err = 0;
  uint32_t e0 = 0, ed = 0, f0 = 0, f4 = 0;
  if (VALIDATE) {
    const uint32_t must = fsl(c.l2, l2, 1) | must2 | m4e;
    const uint32_t n321 = ~(B[3] | B[2] | B[1]);
    const uint32_t is3 = l3 & ~B[4];
    e0 = is3 & n321 & ~B[0];
    ed = is3 & B[3] & B[2] & ~B[1] & B[0];
    f0 = l4 & n321 & ~B[0];
    f4 = l4 & ~B[3] & B[2] & ~B[1] & ~B[0];
    const uint32_t over2 = l2 & ~B[5] & ~B[4] & n321;
    const uint32_t big = l4 & (B[3] | (B[2] & (B[1] | B[0])));
    const uint32_t pe0 = fsl(c.e0, e0, 1), ped = fsl(c.ed, ed, 1), pf0 = fsl(c.f0, f0, 1), pf4 = fsl(c.f4, f4, 1);
    const uint32_t rng = (pe0 & ~B[5]) | (ped & B[5]) | (pf0 & ~B[5] & ~B[4]) | (pf4 & (B[5] | B[4]));
    err = (must ^ cont) | over2 | big | rng;
  }
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int k = 0; k < 6; k++) c.b[k] = B[k];
  c.l2 = l2; c.l3 = l3; c.l4 = l4;
  c.e0 = e0; c.ed = ed; c.f0 = f0; c.f4 = f4;
  return err;
}

// Validation detector alone (validate_utf8[_with_errors]): needs no payload planes.
struct VCarry {
  uint32_t l2, l3, l4, e0, ed, f0, f4;
};
B200_HD VCarry vcarry_from_word(uint32_t pw) {
  const Carry c = carry_from_word(pw);
  VCarry v;
  v.l2 = c.l2; v.l3 = c.l3; v.l4 = c.l4; v.e0 = c.e0; v.ed = c.ed; v.f0 = c.f0; v.f4 = c.f4;
  return v;
}
B200_HD uint32_t utf8_check_block(const uint32_t (&B)[8], VCarry &c) {
  const uint32_t b7 = B[7];
  const uint32_t cont = b7 & ~B[6];
  const uint32_t l2 = b7 & B[6], l3 = l2 & B[5], l4 = l3 & B[4];
  const uint32_t must = fsl(c.l2, l2, 1) | fsl(c.l3, l3, 2) | fsl(c.l4, l4, 3);
  const uint32_t n321 = ~(B[3] | B[2] | B[1]);
  const uint32_t is3 = l3 & ~B[4];
  const uint32_t e0 = is3 & n321 & ~B[0];
  const uint32_t ed = is3 & B[3] & B[2] & ~B[1] & B[0];
  const uint32_t f0 = l4 & n321 & ~B[0];
  const uint32_t f4 = l4 & ~B[3] & B[2] & ~B[1] & ~B[0];
  const uint32_t over2 = l2 & ~B[5] & ~B[4] & n321;
  const uint32_t big = l4 & (B[3] | (B[2] & (B[1] | B[0])));
  const uint32_t pe0 = fsl(c.e0, e0, 1), ped = fsl(c.ed, ed, 1), pf0 = fsl(c.f0, f0, 1), pf4 = fsl(c.f4, f4, 1);
  const uint32_t rng = (pe0 & ~B[5]) | (ped & B[5]) | (pf0 & ~B[5] & ~B[4]) | (pf4 & (B[5] | B[4]));
  c.l2 = l2; c.l3 = l3; c.l4 = l4; c.e0 = e0; c.ed = ed; c.f0 = f0; c.f4 = f4;
  return (must ^ cont) | over2 | big | rng;
}

// ---------------------------------------------------------------------------------------------
// UTF-16LE -> UTF-8 (reference src/scalar/utf16_to_utf8/utf16_to_utf8.h:82-153; surrogate rule src/scalar/utf16.h:39-67).
// Planes W[0..15] of 32 units in split order.  Every unit emits 1..3 bytes on its own: ASCII 1, < 0x800 2, other
// BMP 3, and each half of a surrogate pair 2 (the high surrogate the first two bytes of the 4-byte sequence, the low
// surrogate the last two, which need the two low bits of the high surrogate: one unit of look-back).
// X[0..23] = planes of (byte0 | byte1 << 8 | byte2 << 16) per unit; e1 / e2 = units that emit a second / third byte.
// The detector plane is set where "previous unit is a high surrogate" != "this unit is a low surrogate".
// ---------------------------------------------------------------------------------------------
struct Carry16 {
  uint32_t his, w0, w1;  // previous block: high-surrogate plane and planes 0, 1 (only bit 31 = its last unit is used)
};
B200_HD Carry16 carry16_from_unit(uint32_t pu) {
  Carry16 c;
  c.his = ((pu & 0xFC00u) == 0xD800u) ? 0x80000000u : 0u;
  c.w0 = (pu & 1u) << 31;
  c.w1 = (pu & 2u) << 30;
  return c;
}
// e1 / e2 alone (units that emit a second / a third byte): what the counting pass of the single-pass transcoder needs.
B200_HD void utf16_emit_masks(const uint32_t (&W)[16], uint32_t &e1, uint32_t &e2) {
  const uint32_t ge800 = W[11] | W[12] | W[13] | W[14] | W[15];
  const uint32_t sur = W[15] & W[14] & ~W[13] & W[12] & W[11];
  e1 = ge800 | W[7] | W[8] | W[9] | W[10];
  e2 = ge800 & ~sur;
}
B200_HD uint32_t utf16_to_utf8_block(const uint32_t (&W)[16], Carry16 &c, uint32_t (&X)[32], uint32_t &e1, uint32_t &e2) {
  const uint32_t ge800 = W[11] | W[12] | W[13] | W[14] | W[15];
  const uint32_t na = ge800 | W[7] | W[8] | W[9] | W[10];                // not ASCII
  const uint32_t sur = W[15] & W[14] & ~W[13] & W[12] & W[11];
  const uint32_t his = sur & ~W[10], los = sur & W[10];
  const uint32_t three = ge800 & ~sur;
  const uint32_t two = na & ~ge800;
  const uint32_t p0 = split_prev1(c.w0, W[0]), p1 = split_prev1(c.w1, W[1]);
  const uint32_t phis = split_prev1(c.his, his);
  // t = (high surrogate & 0x3FF) + 0x40 = code point >> 10; bits 0..5 are W[0..5]
  const uint32_t t6 = ~W[6];
  const uint32_t c6 = W[6];
  const uint32_t t7 = W[7] ^ c6, c7 = W[7] & c6;
  const uint32_t t8 = W[8] ^ c7, c8 = W[8] & c7;
  const uint32_t t9 = W[9] ^ c8, c9 = W[9] & c8;
  const uint32_t t10 = c9;
  // first byte
  X[7] = na;
  X[6] = (~na & W[6]) | (na & ~los);
  X[5] = (~na & W[5]) | three | his | (los & p1);
  X[4] = (~na & W[4]) | (two & W[10]) | his | (los & p0);
  X[3] = (~na & W[3]) | (two & W[9]) | (three & W[15]) | (los & W[9]);
  X[2] = (~na & W[2]) | (two & W[8]) | (three & W[14]) | (his & t10) | (los & W[8]);
  X[1] = (~na & W[1]) | (two & W[7]) | (three & W[13]) | (his & t9) | (los & W[7]);
  X[0] = (~na & W[0]) | (two & W[6]) | (three & W[12]) | (his & t8) | (los & W[6]);
  // second byte: 10xxxxxx with x = W[11..6] (three), t[7..2] (high surrogate), W[5..0] (two, low surrogate)
  X[15] = 0xFFFFFFFFu;
  X[14] = 0u;
  X[13] = (three & W[11]) | (his & t7) | (~three & ~his & W[5]);
  X[12] = (three & W[10]) | (his & t6) | (~three & ~his & W[4]);
  X[11] = (three & W[9]) | (his & W[5]) | (~three & ~his & W[3]);
  X[10] = (three & W[8]) | (his & W[4]) | (~three & ~his & W[2]);
  X[9] = (three & W[7]) | (his & W[3]) | (~three & ~his & W[1]);
  X[8] = (three & W[6]) | (his & W[2]) | (~three & ~his & W[0]);
  // third byte: 10 W[5..0]
  X[23] = 0xFFFFFFFFu;
  X[22] = 0u;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int k = 0; k < 6; k++) X[16 + k] = W[k];
  e1 = na;
  e2 = three;
  const uint32_t err = phis ^ los;
  c.his = his;
  c.w0 = W[0];
  c.w1 = W[1];
  return err;
}

// ---------------------------------------------------------------------------------------------
// base64 (reference src/tables/base64_tables.h:791-849 for the classes, src/scalar/base64.h:33-216 for the decode):
// planes B[0..7] of 32 characters -> `valid` (the character is a sextet of the selected alphabet), `ws` (ASCII
// whitespace ' ' \t \n \f \r), and the six planes of the sextet value.  Everything else is an invalid character.
//   plus_ok / slash_ok select the default alphabet's "+" "/" ; minus_ok / under_ok the URL alphabet's "-" "_"
//   (all-ones or zero words, uniform for the call; base64_default_or_url sets all four).
// ---------------------------------------------------------------------------------------------
struct B64Class {
  uint32_t valid, ws;
};
template <bool VALUES>
B200_HD B64Class base64_classify(const uint32_t (&B)[8], uint32_t plus_ok, uint32_t slash_ok, uint32_t minus_ok,
                                 uint32_t under_ok, uint32_t (&S)[6]) {
  const uint32_t b0 = B[0], b1 = B[1], b2 = B[2], b3 = B[3], b4 = B[4], b5 = B[5], b6 = B[6], b7 = B[7];
  const uint32_t lo5_nz = b4 | b3 | b2 | b1 | b0;
  const uint32_t lo5_ge27 = b4 & b3 & (b2 | (b1 & b0));
  const uint32_t letter = ~b7 & b6 & lo5_nz & ~lo5_ge27;         // 41..5A, 61..7A
  const uint32_t row2 = ~b7 & ~b6 & b5;                          // 20..3F
  const uint32_t digit = row2 & b4 & ~(b3 & (b2 | b1));          // 30..39
  const uint32_t p2 = row2 & ~b4 & b3 & b0;                      // 29, 2B, 2D, 2F
  const uint32_t plus = p2 & ~b2 & b1;                           // 2B
  const uint32_t slash = p2 & b2 & b1;                           // 2F
  const uint32_t minus = p2 & b2 & ~b1;                          // 2D
  const uint32_t under = ~b7 & b6 & ~b5 & b4 & b3 & b2 & b1 & b0;  // 5F
  const uint32_t space = row2 & ~(b4 | b3 | b2 | b1 | b0);       // 20
  const uint32_t ctl = ~(b7 | b6 | b5 | b4) & b3;                // 08..0F
  const uint32_t wsc = ctl & ((~b2 & (b1 ^ b0)) | (b2 & ~b1));   // 09, 0A, 0C, 0D
  const uint32_t c62 = (plus & plus_ok) | (minus & minus_ok);
  const uint32_t c63 = (slash & slash_ok) | (under & under_ok);
  B64Class c;
  c.valid = letter | digit | c62 | c63;
  c.ws = space | wsc;
  if (VALUES) {
    const uint32_t upper = letter & ~b5, lower = letter & b5;
    // d = low5 - 1 (upper: 0..25)
    const uint32_t d0 = ~b0, r0 = ~b0;
    const uint32_t d1 = b1 ^ r0, r1 = ~b1 & r0;
    const uint32_t d2 = b2 ^ r1, r2 = ~b2 & r1;
    const uint32_t d3 = b3 ^ r2, r3 = ~b3 & r2;
    const uint32_t d4 = b4 ^ r3;
    // s = d + 26 (lower: 26..51); 26 = 011010b
    const uint32_t s0 = d0;
    const uint32_t s1 = ~d1, k1 = d1;
    const uint32_t s2 = d2 ^ k1, k2 = d2 & k1;
    const uint32_t s3 = ~(d3 ^ k2), k3 = d3 | k2;
    const uint32_t s4 = ~(d4 ^ k3), k4 = d4 | k3;
    const uint32_t s5 = k4;
    // g = low4 + 52 (digit: 52..61): 110100b + n, n <= 9
    const uint32_t g2 = ~b2, g3 = b3 | b2;
    const uint32_t hi = digit | c62 | c63;  // values >= 52 have bits 5 and 4 set
    S[0] = (upper & d0) | (lower & s0) | (digit & b0) | c63;
    S[1] = (upper & d1) | (lower & s1) | (digit & b1) | c62 | c63;
    S[2] = (upper & d2) | (lower & s2) | (digit & g2) | c62 | c63;
    S[3] = (upper & d3) | (lower & s3) | (digit & g3) | c62 | c63;
    S[4] = (upper & d4) | (lower & s4) | hi;
    S[5] = (lower & s5) | hi;
  }
  return c;
}

// 8 planes (u[k] = plane k of a byte per position; planes that are zero cost the same) -> 32 bytes, in place:
// afterwards u[i] = bytes 4i..4i+3.  The exact inverse of transpose_in.
B200_HD void transpose_out8(uint32_t (&u)[8]) {
  uint32_t e0 = u[0], e1 = u[1], e2 = u[2], e3 = u[3];
  uint32_t o0 = u[4], o1 = u[5], o2 = u[6], o3 = u[7];
  dswap<1, 0x55555555u>(e0, e1);
  dswap<1, 0x55555555u>(e2, e3);
  dswap<1, 0x55555555u>(o0, o1);
  dswap<1, 0x55555555u>(o2, o3);
  dswap<2, 0x33333333u>(e0, e2);
  dswap<2, 0x33333333u>(e1, e3);
  dswap<2, 0x33333333u>(o0, o2);
  dswap<2, 0x33333333u>(o1, o3);
  dswap<4, 0x0F0F0F0Fu>(e0, o0);
  dswap<4, 0x0F0F0F0Fu>(e1, o1);
  dswap<4, 0x0F0F0F0Fu>(e2, o2);
  dswap<4, 0x0F0F0F0Fu>(e3, o3);
  tr4x4(e0, e1, e2, e3);
  tr4x4(o0, o1, o2, o3);
  u[0] = e0; u[2] = e1; u[4] = e2; u[6] = e3;
  u[1] = o0; u[3] = o1; u[5] = o2; u[7] = o3;
}

}  // namespace bp
}  // namespace b200
