// k_base64.cu — K7: WHATWG forgiving-base64 decode on sm_100a, and binary_to_base64
//   implementation::base64_to_binary[_details] for `char` input, every base64_options /
//   last_chunk_handling_options value (reference src/generic/base64.h:40-246, src/scalar/base64.h:33-216,
//   src/tables/base64_tables.h:791-849).
//
// ONE launch on the single-pass skeleton of the transcoders (sp_device.cuh); the header of k_b64_decode_v3 has the
// details.  Characters are classified in bit-plane form (bitplane.h: base64_classify, ~45 bitwise instructions per 32
// characters for the classes + ~45 for the six sextet planes, instead of a table lookup per character).  The first
// invalid character is an atomicMin on its index; the CTA that finishes last strips the trailing whitespace / '='
// (block-cooperative backward scan), fetches the last <= 3 sextet characters and applies the last-chunk and padding
// rules (swar.h:b64_finish).
#include "bitplane.h"
#include "bp_device.cuh"
#include "device_common.cuh"
#include "launch.h"
#include "sp_device.cuh"

namespace b200 {

namespace {

// Bit p set iff byte b0 + p lies inside the buffer.
__device__ __forceinline__ uint32_t range_mask32(const InView &in, unsigned long long b0) {
  long long lo = (long long)in.vbeg - (long long)b0, hi = (long long)in.vend - (long long)b0;
  lo = lo < 0 ? 0 : (lo > 32 ? 32 : lo);
  hi = hi < 0 ? 0 : (hi > 32 ? 32 : hi);
  const uint32_t mhi = hi >= 32 ? 0xFFFFFFFFu : ((1u << (unsigned)hi) - 1u);
  const uint32_t mlo = lo >= 32 ? 0xFFFFFFFFu : ((1u << (unsigned)lo) - 1u);
  return mhi & ~mlo;
}

struct B64Opts {
  uint32_t plus_ok, slash_ok, minus_ok, under_ok;  // all-ones / zero words
};

// ---------------------------------------------------------------------------------------------
// K7 single pass (round 2): the skeleton of k_utf8_transcode_v3 / k_utf16_to_utf8_v3 (sp_device.cuh) with base64
// characters in and bytes out.  ONE launch, the text crosses HBM once (round 1 shipped a counting kernel and a decoding
// kernel that read the text twice and classified it twice).  A worker loads its 32K contiguous characters (256-bit loads), transposes them, classifies them ONCE
// (valid mask, whitespace mask, the six sextet planes), hands the warp's sextet count to the scan warp — the worker that
// delivers the CTA's last total publishes the aggregate and reserves the next tile — transposes the sextet planes back
// and compacts the sextets into the warp's staging buffer i & 1 at alignment ZERO.  Tile i leaves two tiles later, when
// the look-back has long delivered its global sextet rank `goff`: only then is the tile's phase in the 4-sextet quanta
// known (pad = goff & 3 sextets of its first quantum precede the tile; lane 0 fetches them by scanning the input
// backwards, whitespace skipped), the quanta are packed 4 sextets -> 3 bytes into the warp's output staging region and
// streamed out as 16-byte vectors funnel-shifted to the destination's alignment.  A quantum belongs to the tile that holds its last sextet;
// the 1 or 2 bytes of a trailing partial quantum are written by the epilogue, which fetches the stream's last <= 3
// sextets for the last-chunk rules anyway.  A tile without whitespace (interior, every character a sextet) stores its
// sextets with two 128-bit shared stores per block instead of 32 predicated byte stores.
// ---------------------------------------------------------------------------------------------
template <int K, int NW>
struct GeomB64 {
  static constexpr uint32_t kRegionBytes = 32u * K;
  static constexpr uint32_t kTileChars = 32u * kRegionBytes;
  static constexpr uint32_t kCtaTileBytes = (uint32_t)NW * kTileChars;
  static constexpr uint32_t kSxBytes = 16u + kTileChars + 16u;            // <= 3 carried sextets in front, slack behind
  static constexpr uint32_t kOutBytes = kTileChars / 4u * 3u + 48u;       // <= 15 of alignment + the bytes + slack
  static constexpr uint32_t kWarpBytes = 2u * kSxBytes + kOutBytes;
  static constexpr uint32_t kSmemBytes = (uint32_t)NW * kWarpBytes;
  static constexpr int kThreads = (NW + 1) * 32;
};

struct B64Shared {
  uint8_t lut[256];
  unsigned long long found;  // block_find_last: 1 + index, 0 = none
  int is_last;
};

__device__ __forceinline__ InView make_view32(const void *p, size_t len_bytes) {  // 32-byte-aligned base
  InView v;
  const uintptr_t a = reinterpret_cast<uintptr_t>(p);
  v.base = reinterpret_cast<const uint4 *>(a & ~uintptr_t(31));
  v.vbeg = a & 31u;
  v.vend = v.vbeg + len_bytes;
  return v;
}

struct PendingB64 {
  uint32_t wtot = 0, iter = 0, tile = 0;
  bool valid = false;
};

// Largest j in [0, end) whose class satisfies `sextet_only ? class <= 63 : class != 64` (64 = whitespace); -1 if none.
// The whole CTA takes part.
template <class SM>
__device__ long long block_find_last_t(const uint8_t *p, long long end, bool sextet_only, SM &sm) {
  constexpr long long kPer = 16;
  for (long long hi = end; hi > 0; hi -= (long long)blockDim.x * kPer) {
    __syncthreads();
    if (threadIdx.x == 0) sm.found = 0ull;
    __syncthreads();
    long long best = -1;
    const long long lo = hi - (long long)(threadIdx.x + 1) * kPer;  // thread 0 looks at the highest 16 bytes
    for (long long j = lo + kPer - 1; j >= lo && j >= 0; j--) {
      const uint32_t c = sm.lut[p[j]];
      if (sextet_only ? (c <= 63u) : (c != 64u)) { best = j; break; }
    }
    if (best >= 0) atomicMax(&sm.found, (unsigned long long)best + 1ull);
    __syncthreads();
    const long long f = (long long)sm.found - 1;
    if (f >= 0) return f;
  }
  __syncthreads();
  return -1;
}

template <int K, int NW, int MINB>
__global__ void __launch_bounds__((NW + 1) * 32, MINB)
k_b64_decode_v3(const char *ptr, size_t len, uint8_t *out, unsigned long long *desc, uint32_t epoch, uint32_t num_tiles,
                uint32_t num_cta_tiles, Scratch *scr, B64Opts o, uint32_t opt_url, uint32_t opt_both, uint32_t opt_garbage,
                unsigned long long last_chunk, FullResultPOD *res) {
  using Gm = GeomB64<K, NW>;
  extern __shared__ __align__(16) uint32_t smem[];  // [NW] { sextet staging x 2, output staging }
  __shared__ sp::Rings rg;
  __shared__ B64Shared sm;
  static_assert(NW <= 31, "one scan warp lane per worker");
  const InView in = make_view32(ptr, len);
  const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  for (uint32_t i = threadIdx.x; i < 256u; i += blockDim.x) sm.lut[i] = (uint8_t)b64_class(i, opt_url != 0, opt_both != 0);
  if (threadIdx.x == 0) sp::init_rings(rg, NW);
  __syncthreads();

  if (warp == (unsigned)NW) {
    sp::scan_warp<NW, 1>(rg, desc, epoch, num_cta_tiles, scr, nullptr, nullptr);
  } else {
    const uint32_t wbase = (uint32_t)__cvta_generic_to_shared(smem) + warp * Gm::kWarpBytes;
    const uint32_t so_addr = wbase + 2u * Gm::kSxBytes;
    const uint32_t one = (blockDim.x >> 5) - (uint32_t)NW;  // 1, opaque to the assembler (bpd::bump)
    PendingB64 q1, q2;  // tiles i - 1 and i - 2

    // waits for the sextet rank of a pending tile, packs its quanta and streams the bytes out
    auto copy_out = [&](const PendingB64 &q) {
      const unsigned long long goff = sp::wait_goff(rg, q.iter, warp);
      if (q.wtot) {
        const uint32_t sx = wbase + (q.iter & 1u) * Gm::kSxBytes + 16u;  // sextet of rank goff + i at sx + i
        const uint32_t pad = (uint32_t)(goff & 3ull);                   // sextets of the first quantum that precede the tile
        if (pad) {
          // fetch them from the input in front of the tile (whitespace and, in the tolerant modes, garbage skipped): the
          // 16 characters before the tile go to 16 lanes at once — one round trip, where a backward walk by one lane was
          // a chain of dependent loads (ncu: 7 % of the kernel's stall samples) — and the nearest `pad` sextets among
          // them are kept; only a longer run of whitespace is walked
          const uint8_t *p8 = reinterpret_cast<const uint8_t *>(in.base);
          const long long t0q = (long long)q.tile * Gm::kTileChars;
          const long long pos = t0q - 1 - (long long)lane;
          uint32_t cls = 255u;
          if (lane < 16u && pos >= (long long)in.vbeg) cls = sm.lut[p8[pos]];
          const unsigned bm = __ballot_sync(kFull, cls <= 63u);
          const uint32_t rank = (uint32_t)__popc(bm & ((1u << lane) - 1u));
          if (cls <= 63u && rank < pad) bpd::sts_u8(sx - 1u - rank, cls);
          const uint32_t got = (uint32_t)__popc(bm);
          if (got < pad && lane == 0) {
            uint32_t need = pad - got;
            for (long long pos2 = t0q - 17; need && pos2 >= (long long)in.vbeg; pos2--) {
              const uint32_t c2 = sm.lut[p8[pos2]];
              if (c2 <= 63u) {
                bpd::sts_u8(sx - 1u - (pad - need), c2);
                need--;
              }
            }
          }
        }
        __syncwarp();
        const uint32_t have = pad + q.wtot;
        const uint32_t nq = have >> 2;
        const uint32_t nb = 3u * nq;                           // output bytes of this tile
        uint8_t *gdst = out + 3ull * (goff >> 2);              // where they go
        {
          // four quanta per lane and round: 16 sextets (five words, realigned by one byte permute each) -> 12 bytes
          // (three 32-bit stores); staging byte 3 j + k is byte k of quantum j; the stale bytes behind the last whole
          // quantum decode into bytes beyond `nb`, which are never copied out
          const uint32_t ngroups = (nq + 3u) >> 2;
          const uint32_t sel = 0x3210u + 0x1111u * (4u - pad);
          for (uint32_t g = lane; g < ngroups; g += 32u) {
            const uint32_t a = sx + 16u * g;
            const uint32_t x0 = sp::lds_u32(a - 4u);
            uint32_t x1, x2, x3, x4;
            asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(x1), "=r"(x2), "=r"(x3), "=r"(x4) : "r"(a));
            const uint32_t w[4] = {__byte_perm(x0, x1, sel) & 0x3F3F3F3Fu, __byte_perm(x1, x2, sel) & 0x3F3F3F3Fu,
                                   __byte_perm(x2, x3, sel) & 0x3F3F3F3Fu, __byte_perm(x3, x4, sel) & 0x3F3F3F3Fu};
            uint32_t o3[3];
            b64_pack_quanta4(w, o3);  // swar.h (host-tested)
            bpd::sts_u32(so_addr + 12u * g, o3[0]);
            bpd::sts_u32(so_addr + 12u * g + 4u, o3[1]);
            bpd::sts_u32(so_addr + 12u * g + 8u, o3[2]);
          }
        }
        __syncwarp();
        sp::copy_out_bytes_v16(so_addr, nb, gdst, lane);  // the destination's own 16-byte vectors, funnel-shifted out of the staging words
      }
      __syncwarp();  // the staging buffers are about to be rewritten
    };

    for (uint32_t iter = 0;; iter++) {
      const uint32_t slot = iter & 3u;
      const uint32_t ct = sp::wait_ticket(rg, iter);
      if (ct >= num_cta_tiles) {  // CTA-uniform: drain
        if (q2.valid) copy_out(q2);
        if (q1.valid) copy_out(q1);
        break;
      }
      const uint32_t tile = ct * (uint32_t)NW + warp;
      const uint32_t stage_cur = wbase + (iter & 1u) * Gm::kSxBytes + 16u;
      const bool active = tile < num_tiles;
      const unsigned long long t0 = (unsigned long long)tile * Gm::kTileChars;
      const unsigned long long r0 = t0 + (unsigned long long)lane * Gm::kRegionBytes;
      const bool interior = active && t0 >= in.vbeg && t0 + Gm::kTileChars <= in.vend;
      // ---- pass 1: this lane's 32K contiguous characters -> planes -> classes, sextet planes, counts ----
      uint32_t S[K][8], V[K];
      uint32_t cnt = 0, allv = 0xFFFFFFFFu;
#pragma unroll
      for (int j = 0; j < K; j++) {
        uint32_t B[8], Sv[6];
        if (interior) {
          sp::ldg_v8(in.base + (r0 >> 4) + 2 * j, B);
        } else if (active) {
          bool ins;
          load_granule(in, (r0 >> 4) + 2ull * j, &B[0], ins);
          load_granule(in, (r0 >> 4) + 2ull * j + 1ull, &B[4], ins);
        } else {
#pragma unroll
          for (int i = 0; i < 8; i++) B[i] = 0u;
        }
        bp::transpose_in(B);
        const bp::B64Class c = bp::base64_classify<true>(B, o.plus_ok, o.slash_ok, o.minus_ok, o.under_ok, Sv);
        uint32_t v = c.valid, bad = ~(c.valid | c.ws);
        if (!interior) {
          const uint32_t r = active ? range_mask32(in, r0 + 32ull * j) : 0u;
          v &= r;
          bad &= r;
        }
        if (bad && !opt_garbage) {
          const unsigned long long pos = r0 + 32ull * j + (unsigned)(__ffs((int)bad) - 1) - in.vbeg;
          const unsigned long long key = err_key(pos, kInvalidBase64Character);
          if (key < ld_relaxed_u64(&scr->err_key)) report_error(scr, key);
        }
#pragma unroll
        for (int k = 0; k < 6; k++) S[j][k] = Sv[k];
        S[j][6] = 0u;
        S[j][7] = 0u;
        V[j] = v;
        allv &= v;
        cnt += (uint32_t)__popc(v);
      }
      const bool dense = interior && __all_sync(kFull, allv == 0xFFFFFFFFu);  // no whitespace in the whole warp-tile
      const uint32_t incl = bpd::warp_inclusive_u32(cnt);
      const uint32_t wtot = __shfl_sync(kFull, incl, 31);
      const uint32_t excl = incl - cnt;
      uint32_t tn = 0;
      bool took = sp::post_totals<NW>(rg, slot, warp, lane, wtot, ct, desc, epoch, scr, tn);
      // ---- the tile before the previous one leaves its staging buffer, which is this tile's ----
      if (q2.valid) copy_out(q2);
      auto post = [&]() {
        tn = sp::post_ticket(rg, iter + 1u, tn, lane);
        if (tn < num_cta_tiles) {  // pull the next CTA-tile into L2
          const char *nx = reinterpret_cast<const char *>(in.base) + (unsigned long long)tn * Gm::kCtaTileBytes;
#pragma unroll
          for (uint32_t k = 0; k < (Gm::kCtaTileBytes + 4095u) / 4096u; k++) {
            const uint32_t off = k * 4096u + lane * 128u;
            if (off < Gm::kCtaTileBytes) asm volatile("prefetch.global.L2 [%0];" ::"l"(nx + off));
          }
        }
        took = false;
      };
      // ---- pass 2: sextet planes back to one byte per character, compaction at alignment zero ----
      if (dense) {
#pragma unroll
        for (int j = 0; j < K; j++) {
          bp::transpose_out8(S[j]);  // S[j][w] = sextets of characters 4w..4w+3, one per byte
          sp::sts_v4(stage_cur + excl + 32u * j, S[j][0], S[j][1], S[j][2], S[j][3]);
          sp::sts_v4(stage_cur + excl + 32u * j + 16u, S[j][4], S[j][5], S[j][6], S[j][7]);
          if (j == 0 && took) post();
        }
      } else if (active) {
        uint32_t spa = stage_cur + excl;
#pragma unroll
        for (int j = 0; j < K; j++) {
          bp::transpose_out8(S[j]);
          const uint32_t m = V[j];
          uint32_t s[4];
          s[0] = spa;
          s[1] = spa + (uint32_t)__popc(m & 0xFFu);
          s[2] = spa + (uint32_t)__popc(m & 0xFFFFu);
          s[3] = spa + (uint32_t)__popc(m & 0xFFFFFFu);
#pragma unroll
          for (int i = 0; i < 8; i++) {
#pragma unroll
            for (int c = 0; c < 4; c++) {
              const int p = 8 * c + i;
              if (m & (1u << p)) {
                const uint32_t w = S[j][p >> 2];
                const uint32_t b = (p & 3) == 0 ? w : __umulhi(w, 1u << (32 - 8 * (p & 3)));
                bpd::sts_u8(s[c], b);
                s[c] = bpd::bump<1>(s[c], one);
              }
            }
          }
          spa += (uint32_t)__popc(m);
          if (j == 0 && took) post();
        }
      }
      if (took) post();
      __syncwarp();  // the staged sextets are visible to the whole warp
      q2 = q1;
      q1.valid = true; q1.wtot = wtot; q1.iter = iter; q1.tile = tile;
    }
  }

  // ---- epilogue by the CTA that finishes last -----------------------------------------------------
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    sm.is_last = (atomicAdd(&scr->done, 1u) == gridDim.x - 1) ? 1 : 0;
    __threadfence();
  }
  __syncthreads();
  if (!sm.is_last) return;

  const uint8_t *p = reinterpret_cast<const uint8_t *>(ptr);
  unsigned long long srclen = len, equallocation = len;
  uint32_t equalsigns = 0;
  if (!opt_garbage) {  // reference src/generic/base64.h:50-73
    srclen = (unsigned long long)(block_find_last_t(p, (long long)srclen, false, sm) + 1);
    equallocation = srclen;
    if (srclen > 0 && p[srclen - 1] == '=') {
      equallocation = srclen - 1;
      srclen--;
      equalsigns = 1;
      srclen = (unsigned long long)(block_find_last_t(p, (long long)srclen, false, sm) + 1);
      if (srclen > 0 && p[srclen - 1] == '=') {
        equallocation = srclen - 1;
        srclen--;
        equalsigns = 2;
      }
    }
  }
  const unsigned long long key = ld_relaxed_u64(&scr->err_key);
  const unsigned long long V_total = num_cta_tiles ? desc_value(ld_relaxed_u64(desc + (num_cta_tiles - 1u))) : 0ull;
  const bool invalid = key != kNoError && (key >> 8) < srclen;
  uint32_t tail_val[3] = {0, 0, 0};
  uint64_t tail_pos[3] = {0, 0, 0};
  const uint32_t idx = (uint32_t)(V_total & 3ull);
  if (!invalid && srclen > 0) {
    long long end = (long long)srclen;
    for (uint32_t k = 0; k < idx; k++) {
      const long long f = block_find_last_t(p, end, true, sm);
      if (f < 0) break;
      tail_pos[k] = (uint64_t)f;
      tail_val[k] = sm.lut[p[f]];
      end = f;
    }
  }
  if (threadIdx.x == 0) {
    if (invalid) {
      res->error = kInvalidBase64Character;
      res->reserved_ = 0;
      res->input_count = key >> 8;
      res->output_count = 0;  // not pinned by the reference: its own kernels disagree here
    } else {
      // the 1 or 2 bytes of a trailing partial quantum (tail_val[0] is the stream's last sextet)
      uint8_t *tail_out = out + 3ull * (V_total >> 2);
      if (idx == 2u) {
        tail_out[0] = (uint8_t)((tail_val[1] << 2) | (tail_val[0] >> 4));
      } else if (idx == 3u) {
        tail_out[0] = (uint8_t)((tail_val[2] << 2) | (tail_val[1] >> 4));
        tail_out[1] = (uint8_t)((tail_val[1] << 4) | (tail_val[0] >> 2));
      }
      int error;
      uint64_t in_count, out_count;
      b64_finish(srclen, equalsigns, equallocation, V_total, opt_garbage != 0, last_chunk, tail_val, tail_pos, &error,
                 &in_count, &out_count);
      res->error = error;
      res->reserved_ = 0;
      res->input_count = in_count;
      res->output_count = out_count;
    }
    scratch_reset(scr);
  }
}

// ---------------------------------------------------------------------------------------------
// binary_to_base64 (reference include/simdutf/implementation.h:4941-4960, semantics
// src/scalar/base64.h:435-491): a fixed 3 -> 4 map, no scan.  A thread takes 48 input bytes (three 128-bit loads) to
// 64 characters (four 128-bit stores) when both pointers are 16-byte aligned; the value -> character map is
// arithmetic on four sextets per word (+65, +6 from 26, -75 from 52, then the two alphabet-dependent symbols).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t b64_chars4(uint32_t v, uint32_t d62, uint32_t d63) {
  // v: four sextets, one per byte.  (v + k) & 0x80 <=> sextet >= 128 - k; no carries between bytes (v <= 63)
  const uint32_t m26 = ((v + 0x66666666u) >> 7) & 0x01010101u;
  const uint32_t m52 = ((v + 0x4C4C4C4Cu) >> 7) & 0x01010101u;
  const uint32_t m62 = ((v + 0x42424242u) >> 7) & 0x01010101u;
  const uint32_t m63 = ((v + 0x41414141u) >> 7) & 0x01010101u;
  // every intermediate byte stays in 0..255, so the word-wide adds and subtracts never cross a byte
  return v + 0x41414141u + m26 * 6u - m52 * 75u - m62 * d62 + m63 * d63;
}
__device__ __forceinline__ uint32_t b64_spread(uint32_t t) {  // 24-bit group (first byte on top) -> four sextets
  return ((t >> 18) & 0x3Fu) | ((t >> 4) & 0x3F00u) | ((t << 10) & 0x3F0000u) | ((t << 24) & 0x3F000000u);
}
__device__ __forceinline__ uint32_t b64_char1(uint32_t v, uint32_t d62, uint32_t d63) {
  return b64_chars4(v, d62, d63) & 0xFFu;
}

__global__ void __launch_bounds__(kBlock) k_b64_encode(const uint8_t *in, size_t len, uint8_t *out, uint32_t url,
                                                        uint32_t pad) {
  const uint32_t d62 = url ? 13u : 15u, d63 = url ? 49u : 3u;
  const size_t tid = (size_t)blockIdx.x * kBlock + threadIdx.x, nthreads = (size_t)gridDim.x * kBlock;
  const bool aligned = ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) & 15u) == 0;
  const size_t ngroups = aligned ? len / 48 : 0;  // 48 bytes -> 64 characters
  for (size_t g = tid; g < ngroups; g += nthreads) {
    const uint4 *vi = reinterpret_cast<const uint4 *>(in) + 3 * g;
    const uint4 a = ldg_stream_v4(vi), b = ldg_stream_v4(vi + 1), c = ldg_stream_v4(vi + 2);
    const uint32_t w[12] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w, c.x, c.y, c.z, c.w};
    uint32_t o[16];
#pragma unroll
    for (int k = 0; k < 4; k++) {  // 12 bytes (3 words) -> 4 groups -> 16 characters
      const uint32_t w0 = w[3 * k], w1 = w[3 * k + 1], w2 = w[3 * k + 2];
      const uint32_t t0 = __byte_perm(w0, 0u, 0x4012);   // bytes 0 1 2
      const uint32_t t1 = __byte_perm(w0, w1, 0x4345);   // bytes 3 | 0 1 of the next word (top byte: don't care)
      const uint32_t t2 = __byte_perm(w1, w2, 0x4234);   // bytes 2 3 | 0
      const uint32_t t3 = __byte_perm(w2, 0u, 0x4123);   // bytes 1 2 3
      o[4 * k] = b64_chars4(b64_spread(t0), d62, d63);
      o[4 * k + 1] = b64_chars4(b64_spread(t1), d62, d63);
      o[4 * k + 2] = b64_chars4(b64_spread(t2), d62, d63);
      o[4 * k + 3] = b64_chars4(b64_spread(t3), d62, d63);
    }
    uint4 *vo = reinterpret_cast<uint4 *>(out) + 4 * g;
#pragma unroll
    for (int k = 0; k < 4; k++) stg_stream_v4(vo + k, make_uint4(o[4 * k], o[4 * k + 1], o[4 * k + 2], o[4 * k + 3]));
  }
  // whatever the vector path did not cover: whole groups of 3 bytes, one per thread and round
  const size_t done = ngroups * 48, nq = (len - done) / 3;
  for (size_t q = tid; q < nq; q += nthreads) {
    const uint8_t *p = in + done + 3 * q;
    const uint32_t t = ((uint32_t)p[0] << 16) | ((uint32_t)p[1] << 8) | p[2];
    const uint32_t ch = b64_chars4(b64_spread(t), d62, d63);
    uint8_t *d = out + done / 3 * 4 + 4 * q;
    d[0] = (uint8_t)ch; d[1] = (uint8_t)(ch >> 8); d[2] = (uint8_t)(ch >> 16); d[3] = (uint8_t)(ch >> 24);
  }
  if (tid == 0) {  // the 1 or 2 trailing bytes (reference src/scalar/base64.h:466-489)
    const size_t full = len / 3, rem = len - 3 * full;
    const uint8_t *p = in + 3 * full;
    uint8_t *d = out + 4 * full;
    if (rem == 1) {
      const uint32_t t = (uint32_t)p[0] << 16;
      d[0] = (uint8_t)b64_char1((t >> 18) & 63u, d62, d63);
      d[1] = (uint8_t)b64_char1((t >> 12) & 63u, d62, d63);
      if (pad) { d[2] = '='; d[3] = '='; }
    } else if (rem == 2) {
      const uint32_t t = ((uint32_t)p[0] << 16) | ((uint32_t)p[1] << 8);
      d[0] = (uint8_t)b64_char1((t >> 18) & 63u, d62, d63);
      d[1] = (uint8_t)b64_char1((t >> 12) & 63u, d62, d63);
      d[2] = (uint8_t)b64_char1((t >> 6) & 63u, d62, d63);
      if (pad) d[3] = '=';
    }
  }
}

template <int K, int NW, int MINB>
cudaError_t launch_b64_v3(const LaunchCtx &c, const char *in, size_t len, char *out, const B64Opts &o, uint32_t url, uint32_t both,
                          uint32_t garbage, uint64_t last_chunk, void *full_res) {
  using Gm = GeomB64<K, NW>;
  const size_t span = (reinterpret_cast<uintptr_t>(in) & 31u) + len;
  const size_t tiles = (span + Gm::kTileChars - 1) / Gm::kTileChars, cta_tiles = (tiles + NW - 1) / NW;
  if (cta_tiles + 1 > c.desc_capacity || tiles > 0xFFFFFF00ull) return cudaErrorInvalidValue;
  static KernelCache kc;
  int per_sm = 1;
  cudaError_t e = kernel_per_sm(kc, c.device, k_b64_decode_v3<K, NW, MINB>, Gm::kThreads, Gm::kSmemBytes, &per_sm);
  if (e != cudaSuccess) return e;
  const size_t cap = (size_t)c.sm_count * per_sm;
  const unsigned grid = (unsigned)(cta_tiles < cap ? (cta_tiles ? cta_tiles : 1) : cap);
  k_b64_decode_v3<K, NW, MINB><<<grid, Gm::kThreads, Gm::kSmemBytes, c.stream>>>(
      in, len, reinterpret_cast<uint8_t *>(out), c.desc, c.epoch, (uint32_t)tiles, (uint32_t)cta_tiles, c.scratch, o, url, both,
      garbage, last_chunk, static_cast<FullResultPOD *>(full_res));
  count_launch(1);
  return cudaGetLastError();
}

}  // namespace

// Geometry of the shipped kernel: 64 characters per lane (2 KiB warp-tiles), 7 workers + the scan warp per CTA, four
// CTAs per SM.  Measured on B200, ms per GiB of CRLF-wrapped text / of text without whitespace (K = 32-character blocks
// per lane, workers x CTAs per SM):  K=2 7x4 0.893 / 0.734   8x3 0.887 / 0.760   11x2 0.936 / 0.795   16x1 1.055 / 0.939;
// K=3 16x1 0.927 / 0.822;  K=1 7x4 1.113 / 0.958;  round 1's two launches 1.067 / 2.39.  Like UTF-16 -> UTF-8 this
// decoder is bound by its byte stores, which many small CTAs hide best.
constexpr int kB64K = 2, kB64Workers = 7, kB64Ctas = 4;

// Workspace in 8-byte slots: one look-back descriptor per CTA-tile.
size_t base64_tiles(const void *in, size_t len) {
  using Gm = GeomB64<kB64K, kB64Workers>;
  const size_t span = (reinterpret_cast<uintptr_t>(in) & 31u) + len;
  const size_t tiles = (span + Gm::kTileChars - 1) / Gm::kTileChars;
  return (tiles + kB64Workers - 1) / kB64Workers + 2;
}

cudaError_t launch_base64_to_binary(const LaunchCtx &c, const char *in, size_t len, char *out, uint64_t options,
                                    uint64_t last_chunk, void *full_res) {
  // reference include/simdutf/implementation.h:2782-2800 and src/scalar/base64.h:66-69
  const uint32_t url = (options & 1u) ? 1u : 0u;
  const uint32_t both = (options & 8u) ? 1u : 0u;
  const uint32_t garbage = (options == 4u || options == 5u || options == 12u) ? 1u : 0u;
  B64Opts o;
  o.plus_ok = o.slash_ok = (both || !url) ? 0xFFFFFFFFu : 0u;
  o.minus_ok = o.under_ok = (both || url) ? 0xFFFFFFFFu : 0u;
  return launch_b64_v3<kB64K, kB64Workers, kB64Ctas>(c, in, len, out, o, url, both, garbage, last_chunk, full_res);
}

// base64_to_binary for char16_t input (reference include/simdutf/implementation.h:4922-4939; the scalar decoder reads
// `c > 255 ? invalid : table[c]`, src/scalar/base64.h:46-50): the units are narrowed one to one into the stream's
// scratch buffer — a unit above 0xFF becomes 0xFF, which no alphabet accepts — and the byte decoder runs on the copy,
// so positions and counts are the same.
__global__ void __launch_bounds__(kBlock) k_b64_narrow_utf16(const uint16_t *in, size_t len, uint8_t *out) {
  const size_t tid = (size_t)blockIdx.x * kBlock + threadIdx.x, nthreads = (size_t)gridDim.x * kBlock;
  for (size_t i = tid; i < len; i += nthreads) {
    const uint32_t u = in[i];
    out[i] = (uint8_t)(u > 0xFFu ? 0xFFu : u);
  }
}

cudaError_t launch_base64_to_binary_utf16(const LaunchCtx &c, const uint16_t *in, size_t len, char *out, uint64_t options,
                                          uint64_t last_chunk, void *full_res) {
  if (!c.tmp) return cudaErrorInvalidValue;
  const unsigned long long want = (len + kBlock - 1) / kBlock;
  const unsigned long long cap = (unsigned long long)c.sm_count * 32;
  k_b64_narrow_utf16<<<(unsigned)(want < cap ? (want ? want : 1) : cap), kBlock, 0, c.stream>>>(in, len, static_cast<uint8_t *>(c.tmp));
  count_launch(1);
  return launch_base64_to_binary(c, static_cast<const char *>(c.tmp), len, out, options, last_chunk, full_res);
}

// reference src/scalar/base64.h:515-533
size_t base64_length_from_binary(size_t len, uint64_t options) {
  const bool url = (options & 1u) != 0, reverse = (options & 2u) != 0;
  const bool pad = url == reverse;  // default pads, url does not, reverse_padding flips either
  if (!pad) return len / 3 * 4 + ((len % 3) ? (len % 3) + 1 : 0);
  return (len + 2) / 3 * 4;
}

cudaError_t launch_binary_to_base64(const LaunchCtx &c, const char *in, size_t len, char *out, uint64_t options) {
  const bool url = (options & 1u) != 0, reverse = (options & 2u) != 0;
  const unsigned long long want = (len / 48 + kBlock - 1) / kBlock + 1;
  const unsigned long long cap = (unsigned long long)c.sm_count * 16;
  k_b64_encode<<<(unsigned)(want < cap ? want : cap), kBlock, 0, c.stream>>>(
      reinterpret_cast<const uint8_t *>(in), len, reinterpret_cast<uint8_t *>(out), url ? 1u : 0u, url == reverse ? 1u : 0u);
  count_launch(1);
  return cudaGetLastError();
}

}  // namespace b200
