// k_utf8_to_utf16.cu — sm_100a kernels K3a/K3b: convert_utf8_to_utf16le[_with_errors]
// (reference include/simdutf/implementation.h:3709-3745; semantics src/scalar/utf8_to_utf16/utf8_to_utf16.h:128-255).
//
// Two streaming kernels made of independent warps (no CTA barrier in the hot loop, no inter-CTA wait):
//
//   K3a  k_utf16_tile_counts   per warp-tile (32 lanes x K blocks x 32 bytes of contiguous input) the number of
//        UTF-16 units the tile emits, plus one total per chunk of 64 tiles; the CTA that finishes last turns the
//        chunk totals into exclusive chunk offsets (and the grand total).  Byte-SWAR popcounts, HBM-bound.
//   K3b  k_utf8_to_utf16_bp    the transcoder, in BIT-PLANE form (bitplane.h): every lane owns 32*K contiguous
//        bytes, transposes each 32-byte block into 8 bit planes, and then classifies, validates and assembles
//        the 16 planes of the candidate unit of all 32 positions with ~75 bitwise instructions per block (the
//        byte-SWAR transcoder this replaces needed ~70 per FOUR bytes).  The unit planes are transposed back
//        to 16-bit units and compacted (predicated 16-bit shared stores) into the lane's PRIVATE staging region,
//        whose odd word stride keeps the 32 lanes of a store instruction on 32 different banks whatever the
//        text looks like.  The lane's region starts at the unit offset that makes its 16-byte vectors line up
//        with 16-byte-aligned output addresses; a lane fetches the few units in front of its first vector from
//        its left neighbour's region and then streams its own vectors to global memory.
//
// Emission rule (bitplane.h): a unit is emitted at the LAST byte of its character (high surrogates at the third
// byte of a 4-byte sequence), so everything except one "is the next byte a continuation" bit looks backwards.
// K3a counts with exactly the same rule, which is what makes the per-tile offsets exact.
#include <cstdlib>
#include <type_traits>

#include "bitplane.h"
#include "bp_device.cuh"
#include "device_common.cuh"
#include "launch.h"

namespace b200 {

namespace {

using bpd::kChunkTiles;
using bpd::kThreads;
using bpd::kWarpsPerCta;
using bpd::sts_u16;
using bpd::sts_u32;

// W32 = false: UTF-16LE output (16-bit units, 8 per 16-byte vector); W32 = true: UTF-32 (4 per vector).
template <int K, bool W32>
struct Geom {
  static constexpr uint32_t kRegionBytes = 32u * K;              // contiguous input bytes per lane
  static constexpr uint32_t kTileBytes = 32u * kRegionBytes;     // per warp
  static constexpr uint32_t kTileGranules = kTileBytes / 16u;
  static constexpr uint32_t kVec = W32 ? 4u : 8u;                // output elements per 16-byte vector
  static constexpr uint32_t kUnitBytes = W32 ? 4u : 2u;
  // a lane emits at most one element per input byte, in front of which sit up to kVec-1 elements of alignment pad
  static constexpr uint32_t kStrideWords = (((32u * K + kVec) * kUnitBytes) / 4u) | 1u;  // odd: conflict-free lanes
  static constexpr uint32_t kSmemBytes = kWarpsPerCta * 32u * kStrideWords * 4u;
  static constexpr uint32_t kMaxVec = (32u * K + kVec - 1u) / kVec;
};

__device__ __forceinline__ InView make_view16(const void *p, size_t len_bytes) {
  InView v;
  const uintptr_t a = reinterpret_cast<uintptr_t>(p);
  v.base = reinterpret_cast<const uint4 *>(a & ~uintptr_t(15));
  v.vbeg = a & 15u;
  v.vend = v.vbeg + len_bytes;
  return v;
}

__device__ __forceinline__ bool tail_truncated16(const InView &in) {
  const long long e = (long long)in.vend;
  const uint8_t *p = reinterpret_cast<const uint8_t *>(in.base);
  const uint32_t b1 = p[e - 1];
  const uint32_t b2 = (e - 2 >= (long long)in.vbeg) ? p[e - 2] : 0u;
  const uint32_t b3 = (e - 3 >= (long long)in.vbeg) ? p[e - 3] : 0u;
  return u8_incomplete_tail(b1, b2, b3);
}

// A buffer that starts with a continuation byte is invalid at position 0; both kernels then emit nothing
// (bitplane.h explains why the end-of-character rule needs this).
__device__ __forceinline__ bool starts_with_continuation(const InView &in) {
  if (in.vend <= in.vbeg) return false;
  const uint32_t b = __ldg(reinterpret_cast<const uint8_t *>(in.base) + in.vbeg);
  return (b & 0xC0u) == 0x80u;
}

// Packed emit mask of a granule: position p = 4k + b (word k, byte b) lives at bit 8b + 4 + k.
__device__ __forceinline__ uint32_t pack_emit(uint32_t e0, uint32_t e1, uint32_t e2, uint32_t e3) {
  return (e0 >> 3) | (e1 >> 2) | (e2 >> 1) | e3;
}
__device__ __forceinline__ uint32_t packed_inrange(const InView &in, unsigned long long g) {
  const uint32_t m16 = inrange_mask16(in, g);
  uint32_t r = 0;
#pragma unroll
  for (int k = 0; k < 4; k++) r |= unmask4((m16 >> (4 * k)) & 0xFu) >> (3 - k);
  return r;
}

// ---------------------------------------------------------------------------------------------
// K3a: per-tile unit counts (granule layout: lane l, item j owns granule g0 + 32 j + l)
// ---------------------------------------------------------------------------------------------
template <int G>
__device__ __forceinline__ bool granules_interior(const InView &in, unsigned long long g0) {
  const unsigned long long lo = g0 * 16ull, hi = (g0 + 32ull * G) * 16ull;
  return lo >= in.vbeg && hi <= in.vend;
}

template <int G, bool EDGE, bool W32>
__device__ __forceinline__ uint32_t count_tile(const InView &in, unsigned long long g0) {
  const unsigned lane = threadIdx.x & 31u;
  uint32_t w[G][4];
  bool inside[G];
#pragma unroll
  for (int j = 0; j < G; j++) {
    const unsigned long long g = g0 + (unsigned long long)j * 32u + lane;
    if (EDGE) {
      load_granule(in, g, w[j], inside[j]);
    } else {
      const uint4 v = ldg_stream_v4(in.base + g);
      w[j][0] = v.x; w[j][1] = v.y; w[j][2] = v.z; w[j][3] = v.w;
      inside[j] = true;
    }
  }
  uint32_t pw[G], nw[G];
  neighbour_words<G>(in, g0, w, pw, nw);
  uint32_t cnt = 0;
#pragma unroll
  for (int j = 0; j < G; j++) {
    uint32_t nc[5], f[5];
    f[0] = u8_ge_f0(pw[j]);
#pragma unroll
    for (int k = 0; k < 4; k++) {
      nc[k] = u8_noncont(w[j][k]);
      f[k + 1] = u8_ge_f0(w[j][k]);
    }
    nc[4] = u8_noncont(nw[j]);
    uint32_t em[4];
#pragma unroll
    for (int k = 0; k < 4; k++) em[k] = fwd1(nc[k], nc[k + 1]) | (W32 ? 0u : back2(f[k], f[k + 1]));
    uint32_t m = pack_emit(em[0], em[1], em[2], em[3]);
    if (EDGE && !inside[j]) m &= packed_inrange(in, g0 + (unsigned long long)j * 32u + lane);
    cnt += (uint32_t)__popc(m);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(kFull, cnt, o);
  return cnt;
}

// Interior tiles (all bytes and the 16 bytes on either side inside the buffer): the emission count is separable,
//   #positions p in tile with noncont(p + 1)  =  #noncont in tile - noncont(t0) + noncont(t1)
//   #positions p in tile with byte(p - 2) >= 0xF0  =  #(>= 0xF0) in tile + [t0-2] + [t0-1] - [t1-2] - [t1-1]
// so the tile needs two plain popcounts (one packed popc per granule) and a 4-word boundary correction instead of
// shifted masks.  For valid input the two position sets are disjoint and this equals the transcoder's count; for
// invalid input it can only be larger (gaps in an output that is unspecified anyway, never an overlap), and the grand
// total still never exceeds utf16_length_from_utf8 / count_utf8.
template <int G, bool W32>
__device__ __forceinline__ uint32_t count_tile_interior(const InView &in, unsigned long long g0) {
  const unsigned lane = threadIdx.x & 31u;
  uint4 v[G];
#pragma unroll
  for (int j = 0; j < G; j++) v[j] = ldg_stream_v4(in.base + g0 + (unsigned long long)j * 32u + lane);
  const uint32_t *wp = reinterpret_cast<const uint32_t *>(in.base);
  const unsigned long long t0w = g0 * 4ull, t1w = (g0 + 32ull * G) * 4ull;
  const uint32_t wa = __ldg(wp + t0w - 1), wb = __ldg(wp + t0w), wc = __ldg(wp + t1w - 1), wd = __ldg(wp + t1w);
  uint32_t cnt = 0;
#pragma unroll
  for (int j = 0; j < G; j++) {
    const uint32_t w[4] = {v[j].x, v[j].y, v[j].z, v[j].w};
    uint32_t m = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) {
      uint32_t mk = u8_noncont(w[k]);               // bit 7 of every byte
      if (!W32) mk |= u8_ge_f0(w[k]) >> 1;          // bit 6
      m |= mk >> (2 * (3 - k));
    }
    cnt += (uint32_t)__popc(m);
  }
  cnt = bpd::warp_sum_u32(cnt);
  cnt += (u8_noncont(wd) >> 7 & 1u) - (u8_noncont(wb) >> 7 & 1u);
  if (!W32) {
    const uint32_t fa = u8_ge_f0(wa), fc = u8_ge_f0(wc);
    cnt += (fa >> 23 & 1u) + (fa >> 31) - (fc >> 23 & 1u) - (fc >> 31);
  }
  return cnt;
}

template <int G, bool W32>
__global__ void __launch_bounds__(kThreads) k_utf16_tile_counts(const char *ptr, size_t len, uint16_t *tile_cnt,
                                                                 unsigned long long *chunk_off, uint32_t num_tiles,
                                                                 uint32_t num_chunks, Scratch *scr) {
  const InView in = make_view16(ptr, len);
  const bool poison = starts_with_continuation(in);
  bpd::counts_pass(
      [&](uint32_t t) -> uint32_t {
        const unsigned long long g0 = (unsigned long long)t * (32ull * G);
        const unsigned long long lo = g0 * 16ull, hi = (g0 + 32ull * G) * 16ull;
        const bool deep = lo >= in.vbeg + 16ull && hi + 16ull <= in.vend;
        const uint32_t c = deep ? count_tile_interior<G, W32>(in, g0) : count_tile<G, true, W32>(in, g0);
        return poison ? 0u : c;
      },
      tile_cnt, chunk_off, num_tiles, num_chunks, scr);
}

// ---------------------------------------------------------------------------------------------
// K3b: bit-plane transcoder
// ---------------------------------------------------------------------------------------------
// Bit p set iff byte b0 + p lies inside the buffer.
__device__ __forceinline__ uint32_t range_mask32(const InView &in, unsigned long long b0) {
  long long lo = (long long)in.vbeg - (long long)b0, hi = (long long)in.vend - (long long)b0;
  lo = lo < 0 ? 0 : (lo > 32 ? 32 : lo);
  hi = hi < 0 ? 0 : (hi > 32 ? 32 : hi);
  const uint32_t mhi = hi >= 32 ? 0xFFFFFFFFu : ((1u << (unsigned)hi) - 1u);
  const uint32_t mlo = lo >= 32 ? 0xFFFFFFFFu : ((1u << (unsigned)lo) - 1u);
  return mhi & ~mlo;
}

// One warp-tile: 32 lanes x K blocks x 32 bytes of input at virtual byte offset tile * kTileBytes.  The tile's output
// starts at element index coff + (warp sum of before_partial).  Must be called by all 32 lanes.
// BE (UTF-16 only): big-endian units — the low and high byte planes of the unit trade places before the transposition
// back (free), the ASCII paths put the byte into the upper half.
template <int K, bool W32, bool BE = false>
__device__ __forceinline__ void transcode_tile(const InView &in, typename std::conditional<W32, uint32_t, uint16_t>::type *out,
                                               unsigned long long out_units, uint32_t tile, uint32_t before_partial,
                                               unsigned long long coff, uint32_t *region_w, bool poison, uint32_t one,
                                               Scratch *scr) {
  using Gm = Geom<K, W32>;
  using OutT = typename std::conditional<W32, uint32_t, uint16_t>::type;
  const unsigned lane = threadIdx.x & 31u;
  OutT *region = reinterpret_cast<OutT *>(region_w);
  const unsigned long long t0 = (unsigned long long)tile * Gm::kTileBytes;  // virtual byte offsets from in.base
  const unsigned long long r0 = t0 + (unsigned long long)lane * Gm::kRegionBytes;
  const bool interior = t0 >= in.vbeg + 16ull && t0 + Gm::kTileBytes + 16ull <= in.vend;
  // ---- this lane's 32K contiguous bytes, the word before them and the byte after them ----
  uint32_t B[K][8];
  uint32_t pw, nbyte;
  if (interior) {
    const uint4 *gp = in.base + (r0 >> 4);
#pragma unroll
    for (int j = 0; j < K; j++) {
      const uint4 v0 = __ldg(gp + 2 * j), v1 = __ldg(gp + 2 * j + 1);
      B[j][0] = v0.x; B[j][1] = v0.y; B[j][2] = v0.z; B[j][3] = v0.w;
      B[j][4] = v1.x; B[j][5] = v1.y; B[j][6] = v1.z; B[j][7] = v1.w;
    }
    pw = __ldg(reinterpret_cast<const uint32_t *>(in.base) + (r0 >> 2) - 1);
    nbyte = __ldg(reinterpret_cast<const uint8_t *>(in.base) + r0 + Gm::kRegionBytes);
  } else {
#pragma unroll
    for (int j = 0; j < K; j++) {
      bool ins;
      load_granule(in, (r0 >> 4) + 2ull * j, &B[j][0], ins);
      load_granule(in, (r0 >> 4) + 2ull * j + 1ull, &B[j][4], ins);
    }
    pw = load_word_guarded(in, (long long)(r0 >> 2) - 1);
    const unsigned long long np = r0 + Gm::kRegionBytes;
    nbyte = (np >= in.vbeg && np < in.vend) ? (uint32_t)__ldg(reinterpret_cast<const uint8_t *>(in.base) + np) : 0u;
  }
  const unsigned long long goff = coff + bpd::warp_sum_u32(before_partial);

  uint32_t next_nc[K];
#pragma unroll
  for (int j = 0; j < K; j++) {
    const uint32_t nb = (j + 1 < K) ? B[(j + 1 < K) ? j + 1 : j][0] : nbyte;
    next_nc[j] = ((nb & 0xC0u) != 0x80u) ? 1u : 0u;
  }
  uint32_t hi = pw;
#pragma unroll
  for (int j = 0; j < K; j++) {
#pragma unroll
    for (int i = 0; i < 8; i++) hi |= B[j][i];
  }
  const bool ascii_tile = !__any_sync(kFull, (hi & kH) != 0u);

  uint32_t em[K];
  uint32_t cnt = 0;
  uint32_t badblocks = 0;
  unsigned long long G;  // global index of this lane's first element
  uint32_t a;            // its offset inside a 16-byte output vector
  auto lane_offsets = [&]() {
    const uint32_t incl = bpd::warp_inclusive_u32(cnt);
    G = goff + (incl - cnt);
    a = (uint32_t)((out_units + G) & (Gm::kVec - 1u));
  };
  // ---- pass 2: units, compaction into the private region ----
  // The running store address lives in a 32-bit shared-space register; it advances through the multiplier
  // (`one` * 2 + address) and the upper unit of a word is extracted with IMAD.HI, so that the compaction costs
  // the ALU pipe nothing but the predicate extraction.
  if (!ascii_tile) {
    // one straight-line block: the shuffles of the lane-offset scan overlap with the unit logic of block 0
    bp::Carry carry;
    // ---- pass 1: planes, emit masks, counts ----
    carry = bp::carry_from_word(pw);
    uint32_t prev_l4 = carry.l4;
#pragma unroll
    for (int j = 0; j < K; j++) {
      bp::transpose_in(B[j]);
      uint32_t m = W32 ? bp::emit32_mask(B[j], next_nc[j]) : bp::emit16_mask(B[j], prev_l4, next_nc[j]);
      prev_l4 = B[j][7] & B[j][6] & B[j][5] & B[j][4];
      if (!interior) m &= range_mask32(in, r0 + 32ull * j);
      if (poison) m = 0;
      em[j] = m;
      cnt += (uint32_t)__popc(m);
    }
    lane_offsets();
    uint32_t spa = (uint32_t)__cvta_generic_to_shared(region + a);
#pragma unroll
    for (int j = 0; j < K; j++) {
      // four independent store chains (positions 0-7, 8-15, 16-23, 24-31): a chain's address register can
      // only advance once the store before it has read it, so one chain alone would serialise the block
      const uint32_t m = em[j];
      constexpr uint32_t kUB = Gm::kUnitBytes;
      uint32_t s0 = spa;
      uint32_t s1 = spa + kUB * (uint32_t)__popc(m & 0xFFu);
      uint32_t s2 = spa + kUB * (uint32_t)__popc(m & 0xFFFFu);
      uint32_t s3 = spa + kUB * (uint32_t)__popc(m & 0xFFFFFFu);
      if (W32) {
        uint32_t C[32];
        const uint32_t err = bp::utf8_to_utf32_block<true>(B[j], carry, C);
        if (err) badblocks |= 1u << j;
        bp::transpose_out21(C);
#pragma unroll
        for (int i = 0; i < 8; i++) {
          if (m & (1u << i)) {
            sts_u32(s0, C[i]);
            s0 = bpd::bump<4>(s0, one);
          }
          if (m & (1u << (8 + i))) {
            sts_u32(s1, C[8 + i]);
            s1 = bpd::bump<4>(s1, one);
          }
          if (m & (1u << (16 + i))) {
            sts_u32(s2, C[16 + i]);
            s2 = bpd::bump<4>(s2, one);
          }
          if (m & (1u << (24 + i))) {
            sts_u32(s3, C[24 + i]);
            s3 = bpd::bump<4>(s3, one);
          }
        }
      } else {
        uint32_t U[16];
        const uint32_t err = bp::utf8_to_utf16_block<true>(B[j], carry, U);
        if (err) badblocks |= 1u << j;
        if (BE) {
#pragma unroll
          for (int k = 0; k < 8; k++) {
            const uint32_t t = U[k];
            U[k] = U[k + 8];
            U[k + 8] = t;
          }
        }
        bp::transpose_out16(U);
#pragma unroll
        for (int i = 0; i < 8; i++) {
          if (m & (1u << i)) {
            sts_u16(s0, U[i]);
            s0 = bpd::bump<2>(s0, one);
          }
          if (m & (1u << (8 + i))) {
            sts_u16(s1, U[8 + i]);
            s1 = bpd::bump<2>(s1, one);
          }
          if (m & (1u << (16 + i))) {
            sts_u16(s2, __umulhi(U[i], 65536u));
            s2 = bpd::bump<2>(s2, one);
          }
          if (m & (1u << (24 + i))) {
            sts_u16(s3, __umulhi(U[8 + i], 65536u));
            s3 = bpd::bump<2>(s3, one);
          }
        }
      }
      spa = s3;
    }
  } else {
#pragma unroll
    for (int j = 0; j < K; j++) {
      uint32_t m = 0xFFFFFFFFu;
      if (!interior) m &= range_mask32(in, r0 + 32ull * j);
      if (poison) m = 0;
      em[j] = m;
      cnt += (uint32_t)__popc(m);
    }
    lane_offsets();
    if (interior && !poison && a == 0u) {
      // every lane emits exactly 32K elements and the tile's output is vector-aligned (a is the same in all lanes:
      // 32K is a multiple of the vector size): widen in registers and store straight to global memory
      uint4 *gv = reinterpret_cast<uint4 *>(out + G);
#pragma unroll
      for (int j = 0; j < K; j++) {
#pragma unroll
        for (int k = 0; k < 8; k++) {
          const uint32_t w = B[j][k];
          if (W32) {
            stg_stream_v4(gv + 8 * j + k, make_uint4(w & 0xFFu, (w >> 8) & 0xFFu, (w >> 16) & 0xFFu, w >> 24));
          } else if ((k & 1) == 0) {
            const uint32_t w1 = B[j][k + 1];
            constexpr uint32_t s01 = BE ? 0x1404u : 0x4140u, s23 = BE ? 0x3424u : 0x4342u;
            stg_stream_v4(gv + 4 * j + (k >> 1), make_uint4(__byte_perm(w, 0u, s01), __byte_perm(w, 0u, s23),
                                                             __byte_perm(w1, 0u, s01), __byte_perm(w1, 0u, s23)));
          }
        }
      }
      return;  // warp-uniform; nothing staged, no errors possible in an all-ASCII interior tile
    }
    uint32_t spa = (uint32_t)__cvta_generic_to_shared(region + a);
#pragma unroll
    for (int j = 0; j < K; j++) {
      const uint32_t m = em[j];
#pragma unroll
      for (int p = 0; p < 32; p++) {
        if (m & (1u << p)) {
          const uint32_t byte = (B[j][p >> 2] >> (8 * (p & 3))) & 0xFFu;
          if (W32) {
            sts_u32(spa, byte);
            spa = bpd::bump<4>(spa, one);
          } else {
            sts_u16(spa, BE ? byte << 8 : byte);
            spa = bpd::bump<2>(spa, one);
          }
        }
      }
    }
  }
  // ---- exact error location (rare): the detector only says "in this block or the 3 bytes before it" ----
  if (!interior) {
#pragma unroll
    for (int j = 0; j < K; j++) {
      const unsigned long long b0 = r0 + 32ull * j;
      if (b0 < in.vend && in.vend <= b0 + 32ull && tail_truncated16(in)) badblocks |= 1u << j;
    }
  }
  if (badblocks) {
#pragma unroll
    for (int j = 0; j < K; j++) {
      const long long b0 = (long long)(r0 + 32ull * j);
      if (badblocks & (1u << j)) u8_locate_error(in, scr, b0 - 3, b0 + 32);
    }
  }
  __syncwarp();

  // ---- staging -> global ----
  {
    OutT *gbase = out + G - a;  // 16-byte aligned
    const uint32_t end = a + cnt;
    if (__all_sync(kFull, cnt >= Gm::kVec)) {
      // every lane owns the 16-byte vectors that hold its elements, except its last partial one (owned by the
      // lane to its right, which copies the elements in front of its own first one from this lane's tail;
      // the source starts on a vector boundary of that region: prev_end = a mod kVec)
      const uint32_t prev_end = __shfl_up_sync(kFull, end, 1);
      const uint32_t vfull = end / Gm::kVec;
      uint32_t v0 = 0;
      if (W32) {
        if (lane > 0) {
          const uint32_t *src = region_w - Gm::kStrideWords + (prev_end - a);
#pragma unroll
          for (uint32_t u = 0; u < 3; u++)
            if (u < a) region_w[u] = src[u];
        }
        if (lane == 0 && a > 0) {  // the tile's first partial vector is shared with the previous tile
#pragma unroll
          for (uint32_t u = 1; u < 4; u++)
            if (u >= a) gbase[u] = region[u];
          v0 = 1;
        }
        if (lane == 31) {  // the tile's last partial vector is shared with the next tile
#pragma unroll
          for (uint32_t u = 0; u < 3; u++) {
            const uint32_t i = vfull * 4u + u;
            if (i < end) gbase[i] = region[i];
          }
        }
      } else {
        if (lane > 0) {
          const uint32_t *src = region_w - Gm::kStrideWords + ((prev_end - a) >> 1);
#pragma unroll
          for (uint32_t u = 0; u < 3; u++)
            if (2u * u + 2u <= a) region_w[u] = src[u];
          if (a & 1u) region[a - 1u] = reinterpret_cast<const OutT *>(src)[a - 1u];
        }
        if (lane == 0 && a > 0) {  // first partial vector of the tile: 2 + 4 + 8 bytes
          uint32_t i = a;
          if (i & 1u) { gbase[i] = region[i]; i++; }
          if (i & 2u) { *reinterpret_cast<uint32_t *>(gbase + i) = region_w[i >> 1]; i += 2u; }
          if (i == 4u) *reinterpret_cast<uint2 *>(gbase + 4) = make_uint2(region_w[2], region_w[3]);
          v0 = 1;
        }
        if (lane == 31) {  // last partial vector of the tile: 8 + 4 + 2 bytes
          const uint32_t r = end & 7u;
          uint32_t i = vfull * 8u;
          if (r & 4u) { *reinterpret_cast<uint2 *>(gbase + i) = make_uint2(region_w[i >> 1], region_w[(i >> 1) + 1u]); i += 4u; }
          if (r & 2u) { *reinterpret_cast<uint32_t *>(gbase + i) = region_w[i >> 1]; i += 2u; }
          if (r & 1u) gbase[i] = region[i];
        }
      }
      // a lane holds at most 32K + kVec - 1 elements: a fixed, fully predicated sequence with immediate offsets
#pragma unroll
      for (uint32_t v = 0; v < Gm::kMaxVec; v++) {
        if (v >= v0 && v < vfull) {
          uint4 x;
          x.x = region_w[4u * v];
          x.y = region_w[4u * v + 1u];
          x.z = region_w[4u * v + 2u];
          x.w = region_w[4u * v + 3u];
          stg_stream_v4(reinterpret_cast<uint4 *>(gbase) + v, x);
        }
      }
    } else {
      // edge tiles and invalid input: element by element
      for (uint32_t i = a; i < end; i++) gbase[i] = region[i];
    }
  }
}

template <int K, int MINB, bool W32, bool BE = false>
__global__ void __launch_bounds__(kThreads, MINB)
k_utf8_transcode_bp(const char *ptr, size_t len, typename std::conditional<W32, uint32_t, uint16_t>::type *out,
                    const uint16_t *tile_cnt, const unsigned long long *chunk_off, uint32_t num_tiles,
                    uint32_t num_chunks, Scratch *scr, ResultPOD *res) {
  using Gm = Geom<K, W32>;
  using OutT = typename std::conditional<W32, uint32_t, uint16_t>::type;
  extern __shared__ __align__(16) uint32_t smem[];
  const InView in = make_view16(ptr, len);
  const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  const uint32_t nwarps = gridDim.x * kWarpsPerCta;
  uint32_t *region_w = smem + (warp * 32u + lane) * Gm::kStrideWords;  // this lane's private staging region
  const bool poison = starts_with_continuation(in);
  const unsigned long long out_units = (unsigned long long)(reinterpret_cast<uintptr_t>(out) / sizeof(OutT));
  const uint32_t one = blockDim.x >> 8;  // 1, but not a constant the assembler can fold (see bpd::bump)

  for (uint32_t tile = blockIdx.x * kWarpsPerCta + warp; tile < num_tiles; tile += nwarps) {
    if (tile + nwarps < num_tiles) {  // pull this warp's next tile into L2 while this one is transcoded
      const char *nx = reinterpret_cast<const char *>(in.base) + (unsigned long long)(tile + nwarps) * Gm::kTileBytes +
                       (unsigned long long)lane * Gm::kRegionBytes;
      asm volatile("prefetch.global.L2 [%0];" ::"l"(nx));
    }
    // where the tile's elements go: chunk offset + counts of the chunk's earlier tiles
    const uint32_t before = bpd::tile_before_partial(tile_cnt, tile);
    const unsigned long long coff = chunk_off[tile / kChunkTiles];
    transcode_tile<K, W32, BE>(in, out, out_units, tile, before, coff, region_w, poison, one, scr);
    __syncwarp();  // the regions are rewritten by the next tile
  }

  if (grid_last_thread(scr)) {
    bpd::write_result_from_key(res, ld_relaxed_u64(&scr->err_key), chunk_off[num_chunks]);
    scratch_reset(scr);
  }
}

// ---------------------------------------------------------------------------------------------
// K3 fused: counts and transcoding in ONE launch.  The counts pass is memory-bound, the transcoder ALU-bound; run
// back to back they add up, run side by side in one persistent grid they overlap.  Work is handed out by an atomic
// ticket in an order that keeps the counts `lead` chunks (of 64 tiles) ahead of the transcoding:
//     ticket < lead                      count chunk `ticket`
//     ticket = lead + 5 k                count chunk lead + k   (nothing once the input is exhausted)
//     ticket = lead + 5 k + 1 + q        transcode quarter q (16 tiles) of chunk k
// A count item writes its 64 tile counts, then chains its chunk total into the running prefix with the decoupled
// look-back over chunk descriptors (device_common.cuh) — off the critical path, it runs `lead` chunks ahead.  A
// transcode item waits until its chunk's inclusive prefix is published (every item it can wait for was handed out
// before it, so the wait is bounded by those items' run time), reads the 64 tile counts and proceeds exactly like
// the two-launch form.  With lead * 128 KiB well inside the 126 MB L2 the transcoder's input is an L2 hit: the
// input crosses HBM once.
// ---------------------------------------------------------------------------------------------
template <int K, int MINB, bool W32>
__global__ void __launch_bounds__(kThreads, MINB)
k_utf8_transcode_fused(const char *ptr, size_t len, typename std::conditional<W32, uint32_t, uint16_t>::type *out,
                       uint16_t *tile_cnt, unsigned long long *desc, uint32_t epoch, uint32_t num_tiles,
                       uint32_t num_chunks, uint32_t lead, Scratch *scr, ResultPOD *res) {
  using Gm = Geom<K, W32>;
  using OutT = typename std::conditional<W32, uint32_t, uint16_t>::type;
  constexpr int G = 2 * K;
  extern __shared__ __align__(16) uint32_t smem[];
  __shared__ uint32_t s_ticket;
  __shared__ uint32_t s_tot[kWarpsPerCta];
  __shared__ unsigned long long s_incl;
  const InView in = make_view16(ptr, len);
  const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  uint32_t *region_w = smem + (warp * 32u + lane) * Gm::kStrideWords;
  const bool poison = starts_with_continuation(in);
  const unsigned long long out_units = (unsigned long long)(reinterpret_cast<uintptr_t>(out) / sizeof(OutT));
  const uint32_t one = blockDim.x >> 8;
  const uint32_t total_tickets = lead + 5u * num_chunks;

  while (true) {
    __syncthreads();  // s_ticket / s_tot / s_incl of the previous item are no longer read
    if (threadIdx.x == 0) s_ticket = atomicAdd(&scr->ticket, 1u);
    __syncthreads();
    const uint32_t t = s_ticket;
    if (t >= total_tickets) break;
    bool is_count;
    uint32_t chunk, quarter = 0;
    if (t < lead) {
      is_count = true;
      chunk = t;
    } else {
      const uint32_t s = t - lead, k = s / 5u, slot = s - 5u * k;
      is_count = slot == 0u;
      chunk = is_count ? lead + k : k;
      quarter = slot - 1u;
    }
    if (is_count) {
      if (chunk >= num_chunks) continue;
      uint32_t mine = 0;
#pragma unroll 1
      for (uint32_t i = 0; i < kChunkTiles / kWarpsPerCta; i++) {
        const uint32_t tile = chunk * kChunkTiles + i * kWarpsPerCta + warp;
        if (tile >= num_tiles) break;
        const unsigned long long g0 = (unsigned long long)tile * (32ull * G);
        const unsigned long long lo = g0 * 16ull, hi = (g0 + 32ull * G) * 16ull;
        const bool deep = lo >= in.vbeg + 16ull && hi + 16ull <= in.vend;
        uint32_t c = deep ? count_tile_interior<G, W32>(in, g0) : count_tile<G, true, W32>(in, g0);
        if (poison) c = 0;
        if (lane == 0) tile_cnt[tile] = (uint16_t)c;
        mine += c;
      }
      if (lane == 0) s_tot[warp] = mine;
      __threadfence();  // the tile counts are visible before the descriptor that announces them
      __syncthreads();
      if (warp == 0) {
        const uint32_t tot = bpd::warp_sum_u32(lane < (unsigned)kWarpsPerCta ? s_tot[lane] : 0u);
        unsigned long long excl;
        uint32_t aux;
        tile_lookback(desc, epoch, chunk, tot, 0u, excl, aux);
      }
    } else {
      if (threadIdx.x == 0) {
        unsigned long long d = ld_relaxed_u64(desc + chunk);
        for (uint32_t spins = 0; desc_epoch(d) != epoch || desc_status(d) != kStatusPrefix; spins++) {
          if (spins > (1u << 24)) {  // cannot happen (see above); never hang the device on a logic error
            report_error(scr, err_key(0, kOther));
            break;
          }
          __nanosleep(64);
          d = ld_relaxed_u64(desc + chunk);
        }
        __threadfence();
        s_incl = desc_value(d);
      }
      __syncthreads();
      const unsigned long long incl = s_incl;  // elements emitted by chunks 0..chunk
      const uint32_t tbase = chunk * kChunkTiles;
      const uint32_t c0 = tbase + lane < num_tiles ? (uint32_t)__ldcg(tile_cnt + tbase + lane) : 0u;
      const uint32_t c1 = tbase + lane + 32u < num_tiles ? (uint32_t)__ldcg(tile_cnt + tbase + lane + 32u) : 0u;
      const unsigned long long coff = incl - bpd::warp_sum_u32(c0 + c1);  // elements emitted by chunks 0..chunk-1
#pragma unroll 1
      for (uint32_t i = 0; i < 2u; i++) {
        const uint32_t in_chunk = quarter * 16u + i * kWarpsPerCta + warp;
        const uint32_t tile = tbase + in_chunk;
        if (tile >= num_tiles) break;
        const uint32_t before = (lane < in_chunk ? c0 : 0u) + (lane + 32u < in_chunk ? c1 : 0u);
        transcode_tile<K, W32>(in, out, out_units, tile, before, coff, region_w, poison, one, scr);
        __syncwarp();
      }
    }
  }

  if (grid_last_thread(scr)) {
    const unsigned long long total = num_chunks ? desc_value(ld_relaxed_u64(desc + (num_chunks - 1u))) : 0ull;
    bpd::write_result_from_key(res, ld_relaxed_u64(&scr->err_key), total);
    scratch_reset(scr);
  }
}

// Tuning knobs for experiments (tools/, profiles/): B200_TUNE_K = 32-byte blocks per lane (1, 2 or 4),
// B200_TUNE_MINB = resident CTAs per SM the kernel is compiled for.
inline int env_int(const char *name, int lo, int hi, int dflt) {
  const char *e = getenv(name);
  if (!e || !*e) return dflt;
  const int v = atoi(e);
  return (v >= lo && v <= hi) ? v : dflt;
}
inline int tuned_k() {
  static int v = env_int("B200_TUNE_K", 1, 4, 2);
  return v == 3 ? 2 : v;
}
inline int tuned_minb16() {
  static int v = env_int("B200_TUNE_MINB", 1, 4, 3);
  return v;
}

inline size_t tiles_for(const void *in, size_t len_bytes, int k) {
  const size_t span = (reinterpret_cast<uintptr_t>(in) & 15u) + len_bytes;
  const size_t per_tile = (size_t)1024 * k;
  return (span + per_tile - 1) / per_tile;
}
inline size_t workspace_slots(size_t tiles) {
  const size_t chunks = (tiles + kChunkTiles - 1) / kChunkTiles;
  return chunks + 1 + (tiles * sizeof(uint16_t) + 7) / 8 + 1;
}

// Measured on B200 (1 GiB mixed): fused 1.31 ms vs two launches 1.10 ms — the counts pass is memory-bound on its own
// but still costs ~200 ALU-pipe instructions per tile, and the transcoder is ALU-bound, so side by side they add up
// after all (and the item hand-out costs barriers).  Kept as an experiment switch (B200_TUNE_FUSED=1), off by default.
inline bool tuned_fused() {
  static int v = env_int("B200_TUNE_FUSED", 0, 1, 0);
  return v != 0;
}
inline uint32_t tuned_lead() {
  static int v = env_int("B200_TUNE_LEAD", 1, 4096, 256);  // chunks of 64 tiles the counts run ahead (256 = 32 MiB)
  return (uint32_t)v;
}

template <int K, int MINB, bool W32, bool BE = false>
cudaError_t launch_bp(const LaunchCtx &c, const char *in, size_t len, void *out, void *res, size_t tiles) {
  using Gm = Geom<K, W32>;
  using OutT = typename std::conditional<W32, uint32_t, uint16_t>::type;
  const size_t chunks = (tiles + kChunkTiles - 1) / kChunkTiles;
  // workspace carved out of the descriptor array: [chunks + 1] u64 (chunk offsets / chunk descriptors), then one u16
  // per tile
  unsigned long long *chunk_off = c.desc;
  uint16_t *tile_cnt = reinterpret_cast<uint16_t *>(c.cnt);
  if (tuned_fused() && !BE) {
    static int per_sm = 0;
    if (per_sm == 0) {
      cudaError_t e = cudaFuncSetAttribute(k_utf8_transcode_fused<K, MINB, W32>,
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Gm::kSmemBytes);
      if (e != cudaSuccess) return e;
      int n = 0;
      e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k_utf8_transcode_fused<K, MINB, W32>, kThreads, Gm::kSmemBytes);
      if (e != cudaSuccess) return e;
      per_sm = n < 1 ? 1 : n;
    }
    const uint32_t lead = (uint32_t)(chunks < tuned_lead() ? chunks : tuned_lead());
    const size_t items = lead + 5 * chunks;
    const size_t cap = (size_t)c.sm_count * per_sm;
    const unsigned grid = (unsigned)(items < cap ? (items ? items : 1) : cap);
    k_utf8_transcode_fused<K, MINB, W32><<<grid, kThreads, Gm::kSmemBytes, c.stream>>>(
        in, len, static_cast<OutT *>(out), tile_cnt, chunk_off, c.epoch, (uint32_t)tiles, (uint32_t)chunks, lead,
        c.scratch, static_cast<ResultPOD *>(res));
    count_launch(1);
    return cudaGetLastError();
  }
  static int per_sm_emit = 0;
  if (per_sm_emit == 0) {
    cudaError_t e = cudaFuncSetAttribute(k_utf8_transcode_bp<K, MINB, W32, BE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)Gm::kSmemBytes);
    if (e != cudaSuccess) return e;
    int n = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k_utf8_transcode_bp<K, MINB, W32, BE>, kThreads, Gm::kSmemBytes);
    if (e != cudaSuccess) return e;
    per_sm_emit = n < 1 ? 1 : n;
  }
  {
    const size_t cap = (size_t)c.sm_count * 8;
    const unsigned grid = (unsigned)(chunks < cap ? chunks : cap);
    k_utf16_tile_counts<2 * K, W32><<<grid, kThreads, 0, c.stream>>>(in, len, tile_cnt, chunk_off, (uint32_t)tiles,
                                                                    (uint32_t)chunks, c.scratch);
  }
  {
    const size_t ctas = (tiles + kWarpsPerCta - 1) / kWarpsPerCta;
    const size_t cap = (size_t)c.sm_count * per_sm_emit;
    const unsigned grid = (unsigned)(ctas < cap ? ctas : cap);
    k_utf8_transcode_bp<K, MINB, W32, BE><<<grid, kThreads, Gm::kSmemBytes, c.stream>>>(
        in, len, static_cast<OutT *>(out), tile_cnt, chunk_off, (uint32_t)tiles, (uint32_t)chunks, c.scratch,
        static_cast<ResultPOD *>(res));
  }
  count_launch(2);
  return cudaGetLastError();
}

}  // namespace

// Workspace, in 8-byte descriptor slots, the two kernels need for an input of `len` bytes.
size_t utf8_to_utf16_tiles(const void *in, size_t len) { return workspace_slots(tiles_for(in, len, tuned_k() < 2 ? tuned_k() : 2)); }
size_t utf8_to_utf32_tiles(const void *in, size_t len) { return workspace_slots(tiles_for(in, len, 2)); }

cudaError_t launch_convert_utf8_to_utf16(const LaunchCtx &c, const char *in, size_t len, uint16_t *out, void *res,
                                         bool big_endian) {
  if (big_endian) {
    const size_t tiles_be = tiles_for(in, len, 2);
    if (workspace_slots(tiles_be) > c.desc_capacity || tiles_be > 0xFFFFFF00ull) return cudaErrorInvalidValue;
    return launch_bp<2, 3, false, true>(c, in, len, out, res, tiles_be);
  }
  const int k = tuned_k();
  const size_t tiles = tiles_for(in, len, k);
  if (workspace_slots(tiles) > c.desc_capacity || tiles > 0xFFFFFF00ull) return cudaErrorInvalidValue;
  const int mb = tuned_minb16();
  switch (k) {
    case 1:
      return launch_bp<1, 3, false>(c, in, len, out, res, tiles);
    case 4:
      return launch_bp<4, 2, false>(c, in, len, out, res, tiles);
    default:
      if (mb <= 2) return launch_bp<2, 2, false>(c, in, len, out, res, tiles);
      return launch_bp<2, 3, false>(c, in, len, out, res, tiles);
  }
}

cudaError_t launch_convert_utf8_to_utf32(const LaunchCtx &c, const char *in, size_t len, uint32_t *out, void *res) {
  const size_t tiles = tiles_for(in, len, 2);
  if (workspace_slots(tiles) > c.desc_capacity || tiles > 0xFFFFFF00ull) return cudaErrorInvalidValue;
  if (tuned_minb16() <= 1) return launch_bp<2, 1, true>(c, in, len, out, res, tiles);
  return launch_bp<2, 2, true>(c, in, len, out, res, tiles);
}

}  // namespace b200
