// k_utf16_to_utf8.cu — sm_100a kernels K6a/K6b: convert_utf16le_to_utf8[_with_errors]
// (reference include/simdutf/implementation.h:4038-4080; semantics src/scalar/utf16_to_utf8/utf16_to_utf8.h:82-153,
// surrogate rule src/scalar/utf16.h:39-67).
//
// ONE launch, the single-pass skeleton of the UTF-8 -> UTF-16 transcoder (k_utf8_to_utf16.cu, sp_device.cuh): every lane
// transposes its 32 contiguous units into 16 planes (bitplane.h: utf16_to_utf8_block), builds the 24 planes of (byte0,
// byte1, byte2) of every unit, transposes them back to one 24-bit word per unit and compacts the 1..3 bytes per unit with
// predicated byte stores into the warp's staging buffer; the output offsets come from a decoupled look-back that runs two
// tiles ahead of the copy-out.  (Round 1: a counts kernel + a transcoder, the input read twice.)
// First error = minimum over flagged blocks of the exact surrogate verdict.
#include <cstdlib>

#include "bitplane.h"
#include "bp_device.cuh"
#include "device_common.cuh"
#include "launch.h"
#include "sp_device.cuh"

namespace b200 {

namespace {


__device__ __forceinline__ uint32_t swap16x2(uint32_t w) { return __byte_perm(w, 0u, 0x2301); }

// Unit at virtual unit index i (from the aligned base), in host order; zero outside the buffer.
// `be`: the buffer holds big-endian units (the UTF-16BE twins of the reference API).
__device__ __forceinline__ uint32_t unit_guarded(const InView &in, long long i, bool be) {
  const long long pos = i * 2;
  if (pos < (long long)in.vbeg || pos >= (long long)in.vend) return 0u;
  const uint32_t u = (uint32_t)__ldg(reinterpret_cast<const uint16_t *>(in.base) + i);
  return be ? ((u >> 8) | ((u & 0xFFu) << 8)) : u;
}

// Exact first bad surrogate among virtual units [lo, hi) (reference src/scalar/utf16.h:44-60).
static __device__ __noinline__ void u16_locate_error(const uint4 *base, unsigned long long vbeg, unsigned long long vend,
                                                     Scratch *scr, long long lo, long long hi, bool be) {
  InView in;
  in.base = base;
  in.vbeg = vbeg;
  in.vend = vend;
  const long long first = (long long)(vbeg >> 1), last = (long long)(vend >> 1);  // vbeg is even: units are 2-byte aligned
  if (lo < first) lo = first;
  if (hi > last) hi = last;
  for (long long i = lo; i < hi; i++) {
    const uint32_t u = unit_guarded(in, i, be);
    if ((u & 0xF800u) != 0xD800u) continue;
    const bool hp = i > first, hn = i + 1 < last;
    if (u16_bad(u, hp ? unit_guarded(in, i - 1, be) : 0u, hp, hn ? unit_guarded(in, i + 1, be) : 0u, hn)) {
      report_error(scr, err_key((unsigned long long)(i - first), kSurrogate));
      return;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// K6b: bit-plane transcoder
// ---------------------------------------------------------------------------------------------
// Bits (split order) of the units of the 64-byte block at virtual byte offset b0 that lie inside the buffer.
__device__ __forceinline__ uint32_t range_mask_split(const InView &in, unsigned long long b0) {
  uint32_t m = 0;
#pragma unroll
  for (int s = 0; s < 32; s++) {
    const unsigned long long pos = b0 + 2ull * s;
    if (pos >= in.vbeg && pos < in.vend) m |= 1u << bp::split_pos(s);
  }
  return m;
}

// Compaction of one block: unit s (stream order) owns bit split_pos(s) of the masks and word split_pos(s) of X.
// Four independent store chains (units 0-7, 8-15, 16-23, 24-31) so that no chain waits for its own previous store.
template <bool ALL>
__device__ __forceinline__ void compact_block(const uint32_t (&X)[32], uint32_t e0, uint32_t e1, uint32_t e2,
                                              uint32_t base, uint32_t one) {
  uint32_t s[4];
  s[0] = base;
#pragma unroll
  for (int c = 1; c < 4; c++) {
    const uint32_t mk = c == 1 ? 0x000F000Fu : c == 2 ? 0x00FF00FFu : 0x0FFF0FFFu;
    const uint32_t n0 = ALL ? 8u * c : (uint32_t)__popc(e0 & mk);
    s[c] = base + n0 + (uint32_t)__popc(e1 & mk) + (uint32_t)__popc(e2 & mk);
  }
#pragma unroll
  for (int i = 0; i < 8; i++) {
#pragma unroll
    for (int c = 0; c < 4; c++) {
      const int p = bp::split_pos(8 * c + i);
      const uint32_t x = X[p];
      if (ALL || (e0 & (1u << p))) {
        bpd::sts_u8(s[c], x);
        s[c] = bpd::bump<1>(s[c], one);
      }
      if (e1 & (1u << p)) {
        bpd::sts_u8(s[c], __umulhi(x, 1u << 24));
        s[c] = bpd::bump<1>(s[c], one);
      }
      if (e2 & (1u << p)) {
        bpd::sts_u8(s[c], __umulhi(x, 1u << 16));
        s[c] = bpd::bump<1>(s[c], one);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// K6 single pass (round 2): the skeleton of k_utf8_transcode_v3 (k_utf8_to_utf16.cu, sp_device.cuh) with UTF-16 in and
// bytes out.  ONE launch, the input crosses HBM once: a persistent grid of CTAs of NW worker warps + one scan warp takes
// CTA-tiles through an atomic ticket; a worker loads its 32K contiguous units (256-bit loads), transposes them, counts
// the bytes they emit (pass 1: utf16_emit_masks), hands the warp total to the scan warp — the worker that delivers the
// CTA's last total publishes the aggregate and reserves the next tile — builds the 24 byte planes, transposes back and
// compacts into the warp's staging buffer i & 1 at alignment ZERO (pass 2), and copies tile i out two tiles later, when
// the look-back has long delivered its global offset: 32-bit words, 128 contiguous bytes per warp instruction,
// realigned to the destination's word grid by one byte permute.
// ---------------------------------------------------------------------------------------------
template <int K, int NW>
struct GeomV3 {
  static constexpr uint32_t kRegionBytes = 64u * K;                      // 32K units per lane
  static constexpr uint32_t kTileBytes = 32u * kRegionBytes;
  static constexpr uint32_t kCtaTileBytes = (uint32_t)NW * kTileBytes;
  static constexpr uint32_t kTileUnits = kTileBytes / 2u;
  static constexpr uint32_t kStageBytes = 3u * kTileUnits + 16u;         // <= 3 bytes per unit, + the word behind the last byte
  static constexpr uint32_t kSmemBytes = (uint32_t)NW * 2u * kStageBytes;
  static constexpr int kThreads = (NW + 1) * 32;
};

__device__ __forceinline__ InView make_view_u16_32(const uint16_t *p, size_t len_units) {  // 32-byte-aligned base
  InView v;
  const uintptr_t a = reinterpret_cast<uintptr_t>(p);
  v.base = reinterpret_cast<const uint4 *>(a & ~uintptr_t(31));
  v.vbeg = a & 31u;
  v.vend = v.vbeg + 2ull * len_units;
  return v;
}

struct PendingU16 {
  uint32_t wtot = 0, iter = 0;
  bool valid = false;
};

template <int K, int NW, bool BE, int MINB = 1>
__global__ void __launch_bounds__((NW + 1) * 32, MINB)
k_utf16_to_utf8_v3(const uint16_t *ptr, size_t len, uint8_t *out, unsigned long long *desc, uint32_t epoch, uint32_t num_tiles,
                   uint32_t num_cta_tiles, Scratch *scr, ResultPOD *res) {
  using Gm = GeomV3<K, NW>;
  extern __shared__ __align__(16) uint32_t smem[];  // [NW][2] staging buffers
  __shared__ sp::Rings rg;
  static_assert(NW <= 31, "one scan warp lane per worker");
  const InView in = make_view_u16_32(ptr, len);
  const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) sp::init_rings(rg, NW);
  __syncthreads();

  if (warp == (unsigned)NW) {
    sp::scan_warp<NW, 1>(rg, desc, epoch, num_cta_tiles, scr, nullptr, nullptr);
  } else {
    const uint32_t stage0 = (uint32_t)__cvta_generic_to_shared(smem) + warp * 2u * Gm::kStageBytes;
    const uint32_t one = (blockDim.x >> 5) - (uint32_t)NW;  // 1, but not a constant the assembler can fold (bpd::bump)
    const long long last_unit = (long long)(in.vend >> 1) - 1;  // virtual index of the buffer's last unit
    PendingU16 q1, q2;  // tiles i - 1 and i - 2
    auto copy_out = [&](const PendingU16 &q) {
      const unsigned long long goff = sp::wait_goff(rg, q.iter, warp);
      if (q.wtot) sp::copy_out_bytes(stage0 + (q.iter & 1u) * Gm::kStageBytes, q.wtot, out + goff, lane);
      __syncwarp();  // the staging buffer is about to be rewritten
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
      const uint32_t stage_cur = stage0 + (iter & 1u) * Gm::kStageBytes;
      const bool active = tile < num_tiles;
      const unsigned long long t0 = (unsigned long long)tile * Gm::kTileBytes;
      const unsigned long long r0 = t0 + (unsigned long long)lane * Gm::kRegionBytes;
      const bool interior = active && t0 >= in.vbeg + 32ull && t0 + Gm::kTileBytes + 2ull <= in.vend;  // never the tile of the last unit
      // ---- this lane's 32K contiguous units and the unit before them ----
      uint32_t W[K][16];
      uint32_t pu = 0;
      if (interior) {
        const uint4 *gp = in.base + (r0 >> 4);
#pragma unroll
        for (int j = 0; j < K; j++) {
          sp::ldg_v8(gp + 4 * j, &W[j][0]);
          sp::ldg_v8(gp + 4 * j + 2, &W[j][8]);
        }
        if (lane == 0) pu = (uint32_t)__ldg(reinterpret_cast<const uint16_t *>(in.base) + (r0 >> 1) - 1);
        const uint32_t up = __shfl_up_sync(kFull, W[K - 1][15], 1);
        if (lane != 0) pu = up >> 16;
        if (BE) pu = ((pu >> 8) | (pu << 8)) & 0xFFFFu;
      } else if (active) {
#pragma unroll
        for (int j = 0; j < K; j++) {
#pragma unroll
          for (int g = 0; g < 4; g++) {
            bool ins;
            load_granule(in, (r0 >> 4) + 4ull * j + (unsigned long long)g, &W[j][4 * g], ins);
          }
        }
        pu = unit_guarded(in, (long long)(r0 >> 1) - 1, BE);
      } else {
#pragma unroll
        for (int j = 0; j < K; j++) {
#pragma unroll
          for (int i = 0; i < 16; i++) W[j][i] = 0u;
        }
      }
      if (BE) {  // host order from here on
#pragma unroll
        for (int j = 0; j < K; j++) {
#pragma unroll
          for (int i = 0; i < 16; i++) W[j][i] = swap16x2(W[j][i]);
        }
      }
      // ---- pass 1: planes, emit masks, counts ----
      uint32_t hi = 0;
#pragma unroll
      for (int j = 0; j < K; j++) {
#pragma unroll
        for (int i = 0; i < 16; i++) hi |= W[j][i];
      }
      hi = (hi & 0xFF80FF80u) | (pu & 0xFF80u);
      const bool ascii = interior && !__any_sync(kFull, hi != 0u);  // every lane emits exactly 32K bytes
      uint32_t e0[K];
      uint32_t cnt = 0;
      if (!ascii) {
#pragma unroll
        for (int j = 0; j < K; j++) {
          bp::transpose_in16(W[j]);
          e0[j] = interior ? 0xFFFFFFFFu : (active ? range_mask_split(in, r0 + 64ull * j) : 0u);
          uint32_t e1, e2;
          bp::utf16_emit_masks(W[j], e1, e2);
          cnt += (uint32_t)__popc(e0[j]) + (uint32_t)__popc(e1 & e0[j]) + (uint32_t)__popc(e2 & e0[j]);
        }
      } else {
        cnt = 32u * K;
      }
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
      // ---- pass 2: byte planes, transposition back, compaction into the staging buffer at alignment zero ----
      if (active && !ascii) {
        bp::Carry16 carry = bp::carry16_from_unit(pu);
        uint32_t spa = stage_cur + excl;
        bool bad = false;
#pragma unroll
        for (int j = 0; j < K; j++) {
          uint32_t X[32];
          uint32_t e1, e2;
          const uint32_t err = bp::utf16_to_utf8_block(W[j], carry, X, e1, e2);
          e1 &= e0[j];
          e2 &= e0[j];
          bp::transpose_out_n<24>(X);
          if (interior) compact_block<true>(X, e0[j], e1, e2, spa, one);
          else compact_block<false>(X, e0[j], e1, e2, spa, one);
          spa += (uint32_t)__popc(e0[j]) + (uint32_t)__popc(e1) + (uint32_t)__popc(e2);
          if (err) {
            const long long u0 = (long long)((r0 + 64ull * j) >> 1);
            u16_locate_error(in.base, in.vbeg, in.vend, scr, u0 - 1, u0 + 32, BE);
          }
          if (j == 0 && took) post();  // the ticket has had a block's worth of time to come back
          (void)bad;
        }
        if (!interior) {  // a high surrogate cut off by the end of the buffer
          const long long u0 = (long long)(r0 >> 1);
          if (last_unit >= u0 && last_unit < u0 + 32 * K && in.vend > in.vbeg) {
            const uint32_t lu = unit_guarded(in, last_unit, BE);
            if ((lu & 0xFC00u) == 0xD800u) u16_locate_error(in.base, in.vbeg, in.vend, scr, last_unit - 1, last_unit + 1, BE);
          }
        }
      } else if (ascii) {
        // 32K units -> 32K bytes per lane, at byte excl = 32K * lane of the staging buffer
#pragma unroll
        for (int j = 0; j < K; j++) {
          uint32_t b[8];
#pragma unroll
          for (int i = 0; i < 8; i++) b[i] = __byte_perm(W[j][2 * i], W[j][2 * i + 1], 0x6420);
          sp::sts_v4(stage_cur + excl + 32u * j, b[0], b[1], b[2], b[3]);
          sp::sts_v4(stage_cur + excl + 32u * j + 16u, b[4], b[5], b[6], b[7]);
        }
      }
      if (took) post();
      __syncwarp();  // the staged bytes are visible to the whole warp
      q2 = q1;
      q1.valid = true; q1.wtot = wtot; q1.iter = iter;
    }
  }

  if (grid_last_thread(scr)) {
    const unsigned long long total = num_cta_tiles ? desc_value(ld_relaxed_u64(desc + (num_cta_tiles - 1u))) : 0ull;
    bpd::write_result_from_key(res, ld_relaxed_u64(&scr->err_key), total);
    scratch_reset(scr);
  }
}

template <int K, int NW, bool BE, int MINB = 1>
cudaError_t launch_u16to8_v3(const LaunchCtx &c, const uint16_t *in, size_t len, char *out, void *res) {
  using Gm = GeomV3<K, NW>;
  const size_t span = (reinterpret_cast<uintptr_t>(in) & 31u) + 2 * len;
  const size_t tiles = (span + Gm::kTileBytes - 1) / Gm::kTileBytes, cta_tiles = (tiles + NW - 1) / NW;
  if (cta_tiles + 1 > c.desc_capacity || tiles > 0xFFFFFF00ull) return cudaErrorInvalidValue;
  static KernelCache kc;
  int per_sm = 1;
  cudaError_t e = kernel_per_sm(kc, c.device, k_utf16_to_utf8_v3<K, NW, BE, MINB>, Gm::kThreads, Gm::kSmemBytes, &per_sm);
  if (e != cudaSuccess) return e;
  const size_t cap = (size_t)c.sm_count * per_sm;
  const unsigned grid = (unsigned)(cta_tiles < cap ? (cta_tiles ? cta_tiles : 1) : cap);
  k_utf16_to_utf8_v3<K, NW, BE, MINB><<<grid, Gm::kThreads, Gm::kSmemBytes, c.stream>>>(
      in, len, reinterpret_cast<uint8_t *>(out), c.desc, c.epoch, (uint32_t)tiles, (uint32_t)cta_tiles, c.scratch,
      static_cast<ResultPOD *>(res));
  count_launch(1);
  return cudaGetLastError();
}

}  // namespace

// Workspace, in 8-byte descriptor slots: one look-back descriptor per CTA-tile.
constexpr int kU16Workers = 7;
size_t utf16_convert_tiles(const void *in, size_t len) {
  const size_t span = (reinterpret_cast<uintptr_t>(in) & 31u) + 2 * len;
  const size_t tiles = (span + GeomV3<1, kU16Workers>::kTileBytes - 1) / GeomV3<1, kU16Workers>::kTileBytes;
  return (tiles + kU16Workers - 1) / kU16Workers + 2;
}

cudaError_t launch_convert_utf16_to_utf8(const LaunchCtx &c, const uint16_t *in, size_t len, char *out, void *res,
                                         bool big_endian) {
  // four CTAs of 7 workers + the scan warp per SM, 64 bytes (32 units) per lane.  Measured on B200, 2 GiB of mixed
  // UTF-16, ms per launch (K = blocks per lane, workers x CTAs): K=2 16x1 2.58; K=1 16x1 2.49, 20x1 2.36, 9x3 2.06,
  // 8x3 2.05, 5x5 2.04, 7x4 1.99, 6x4 1.99; round 1's two launches (counts + transcoder with lane-private staging): 2.24.
  // Unlike UTF-8 -> UTF-16 (one fat CTA per SM), this transcoder is bound by its byte stores (three predicated shared
  // stores per unit), and many small CTAs hide their latency best.
  if (big_endian) return launch_u16to8_v3<1, kU16Workers, true, 4>(c, in, len, out, res);
  return launch_u16to8_v3<1, kU16Workers, false, 4>(c, in, len, out, res);
}

}  // namespace b200
