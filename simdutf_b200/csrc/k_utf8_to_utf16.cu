// k_utf8_to_utf16.cu — sm_100a kernel K3: convert_utf8_to_utf16le/be[_with_errors], convert_utf8_to_utf32[_with_errors]
// (reference include/simdutf/implementation.h:3709-3800; semantics src/scalar/utf8_to_utf16/utf8_to_utf16.h:128-255,
// src/scalar/utf8_to_utf32/utf8_to_utf32.h:106-212).
//
// ONE launch, the input crosses HBM once (round 1 shipped a counts kernel + a transcoding kernel: the text was read
// twice, which capped the roofline fraction at 2/3).  A persistent grid of CTAs of NW worker warps + one scan warp
// takes CTA-tiles (NW warp-tiles of 32 lanes x K blocks x 32 contiguous bytes) in increasing order through an atomic
// ticket.  Per warp-tile:
//
//   pass 1   every lane loads its 32K contiguous bytes (one 256-bit load per block), transposes each 32-byte block into
//            8 BIT PLANES (bitplane.h) and derives the block's emit mask; popcounts -> lane count -> warp inclusive scan
//            -> warp total, handed to the scan warp.
//   pass 2   bp::utf8_to_utf16_block (validation detector + the 16 planes of the candidate unit of all 32
//            positions, ~75 bitwise instructions per block), transposition back to 16-bit units, compaction with
//            predicated 16-bit shared stores into the WARP's staging buffer at alignment zero (lane l starts at the
//            warp prefix of the lane counts): no global offset is needed to get this far.
//   copy-out two tiles later, when the look-back has long delivered the tile's global offset: 32-bit words, 128
//            contiguous bytes per warp instruction, one byte permute per word when the destination starts on an odd
//            unit.
//
// The planes stay in registers from pass 1 to pass 2, so counting costs nothing beyond what the transcoder needs anyway.
// The header of k_utf8_transcode_v3 explains how the chained scan is kept off the workers' path.
//
// Emission rule (bitplane.h): a unit is emitted at the LAST byte of its character (high surrogates at the third
// byte of a 4-byte sequence), so everything except one "is the next byte a continuation" bit looks backwards.
#include <type_traits>

#include "bitplane.h"
#include "bp_device.cuh"
#include "device_common.cuh"
#include "launch.h"
#include "sp_device.cuh"

namespace b200 {

namespace {

using bpd::sts_u16;
using bpd::sts_u32;
using sp::atom_add_u32;
using sp::lds_u16;
using sp::lds_u32;
using sp::mbar_arrive;
using sp::mbar_wait_hint;
using sp::stg_cs_u32;

// The input as 32-byte granules: a 32-byte-aligned base and the half-open range [vbeg, vend) of byte offsets that belong
// to the caller's buffer (the kernel loads 32 bytes per instruction).
__device__ __forceinline__ InView make_view32(const void *p, size_t len_bytes) {
  InView v;
  const uintptr_t a = reinterpret_cast<uintptr_t>(p);
  v.base = reinterpret_cast<const uint4 *>(a & ~uintptr_t(31));
  v.vbeg = a & 31u;
  v.vend = v.vbeg + len_bytes;
  return v;
}
__device__ __forceinline__ void ldg_v8(const void *p, uint32_t (&r)[8]) {
  asm volatile("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "l"(p));
}

__device__ __forceinline__ bool tail_truncated16(const InView &in) {
  const long long e = (long long)in.vend;
  const uint8_t *p = reinterpret_cast<const uint8_t *>(in.base);
  const uint32_t b1 = p[e - 1];
  const uint32_t b2 = (e - 2 >= (long long)in.vbeg) ? p[e - 2] : 0u;
  const uint32_t b3 = (e - 3 >= (long long)in.vbeg) ? p[e - 3] : 0u;
  return u8_incomplete_tail(b1, b2, b3);
}

// A buffer that starts with a continuation byte is invalid at position 0; the kernel then emits nothing
// (bitplane.h explains why the end-of-character rule needs this).
__device__ __forceinline__ bool starts_with_continuation(const InView &in) {
  if (in.vend <= in.vbeg) return false;
  const uint32_t b = __ldg(reinterpret_cast<const uint8_t *>(in.base) + in.vbeg);
  return (b & 0xC0u) == 0x80u;
}

// Bit p set iff byte b0 + p lies inside the buffer.
__device__ __forceinline__ uint32_t range_mask32(const InView &in, unsigned long long b0) {
  long long lo = (long long)in.vbeg - (long long)b0, hi = (long long)in.vend - (long long)b0;
  lo = lo < 0 ? 0 : (lo > 32 ? 32 : lo);
  hi = hi < 0 ? 0 : (hi > 32 ? 32 : hi);
  const uint32_t mhi = hi >= 32 ? 0xFFFFFFFFu : ((1u << (unsigned)hi) - 1u);
  const uint32_t mlo = lo >= 32 ? 0xFFFFFFFFu : ((1u << (unsigned)lo) - 1u);
  return mhi & ~mlo;
}

// The workspace (one descriptor per CTA-tile) is sized for CTA-tiles of at least kWorkers warp-tiles.
constexpr int kWorkers = 7;

// Hand-offs go through mbarrier objects in shared memory (one arrival releases any number of waiters, and waiters do not
// wait for EACH OTHER the way the threads of a bar.sync do): a worker that is ahead never waits for a slower worker.

// ---------------------------------------------------------------------------------------------
// The single-pass transcoder, with the look-back OFF the workers' path.
//
// In this round's first single-pass kernel (k_utf8_transcode_sp, git history) a worker needed the tile's global output
// offset BEFORE it compacted (the staging position carried the alignment of the output address), so the planes were
// parked in shared memory for one tile and every hiccup of the chained scan showed up as a wait in front of pass 2
// (ncu: 24 % of all samples on that one mbarrier).  Here a worker transcodes and compacts tile i into staging buffer
// i & 1 at alignment ZERO, knowing nothing but the warp's own lane prefix; the global offset is needed only by the
// copy-out, and the copy-out of tile i runs TWO tiles later, between pass 1 and pass 2 of tile i + 2 (just before that
// buffer is overwritten): the look-back of a tile has two whole tile times to finish.
//
// Who does what, and why (each step measured on B200 with the clock64 / globaltimer instrumentation of the DBG
// instantiation, tools/dbg_timing.py; 1 GiB of mixed text):
//   * the worker that delivers the LAST warp total of a tile publishes the CTA aggregate itself.  With the scan warp
//     publishing it, an aggregate waited for the scan warp's previous look-back, which waited for other CTAs'
//     aggregates, ...: 13 us and 25 polls per look-back; published by the workers, 5 us and 5 polls;
//   * the same worker reserves the CTA's next tile (global atomic ticket); the ticket's round trip hides behind its
//     copy-out.  Reserved by the FIRST worker to finish, a CTA whose warps had drifted apart held a ticket for up to two
//     tile times before its aggregate appeared: the 32 nearest predecessors of a tile published 7 us AFTER it on
//     average, every look-back waited that long, and the workers stalled 1.2 us per tile on the offsets; reserved by
//     the last one, 2 us and 0.06 us.  Tickets (not blockIdx) keep the scan deadlock-free when not all CTAs are resident;
//   * the scan warp only looks back and posts the offsets.  Its loads are COALESCED (a warp load reads 32 consecutive
//     descriptors = 8 sectors, and a lane re-reads only a descriptor that is not ready): a persistent grid of equal
//     tiles drifts into lockstep, then no predecessor of the current wave has its prefix yet and every CTA reads the
//     whole in-flight window, G descriptors G times per wave, on a few dozen L2 lines — with lane-strided loads
//     widening the window from 128 to 512 descriptors took the kernel from 2.3 to 5.6 ms per GiB.
// Not the bottleneck (measured, so that nobody tries again): the shared-store bank conflicts of the warp-contiguous
// staging buffer (3.5 wavefronts per store on the mix; a conflict-free but wrong placement ran 0.853 instead of 0.867
// ms), hiding the tile's global loads behind the copy-out (0.893), lane-private staging regions with per-lane vector
// stores (1.59: the eight alignment-specialised copy loops overflowed the instruction cache).
// ---------------------------------------------------------------------------------------------
template <int K, bool W32, int NW>
struct Geom3 {
  static constexpr uint32_t kRegionBytes = 32u * K;
  static constexpr uint32_t kTileBytes = 32u * kRegionBytes;
  static constexpr uint32_t kCtaTileBytes = (uint32_t)NW * kTileBytes;
  static constexpr uint32_t kUnitBytes = W32 ? 4u : 2u;
  static constexpr uint32_t kStageBytes = kTileBytes * kUnitBytes + 16u;  // + the word behind the last unit (odd starts)
  static constexpr uint32_t kSmemBytes = (uint32_t)NW * 2u * kStageBytes;
  static constexpr int kThreads = (NW + 1) * 32;
};


// staging (elements [0, n) at alignment zero) -> dst[0 .. n)
template <bool W32, class OutT>
__device__ __forceinline__ void copy_out_staged(uint32_t stage_addr, uint32_t n, OutT *dst, unsigned lane) {
  if (W32) {
#pragma unroll 4
    for (uint32_t w = lane; w < n; w += 32u) stg_cs_u32(dst + w, lds_u32(stage_addr + 4u * w));
  } else {
    if ((reinterpret_cast<uintptr_t>(dst) & 2u) == 0u) {
      uint32_t *d32 = reinterpret_cast<uint32_t *>(dst);
      const uint32_t nw = n >> 1;
#pragma unroll 4
      for (uint32_t w = lane; w < nw; w += 32u) stg_cs_u32(d32 + w, lds_u32(stage_addr + 4u * w));
      if ((n & 1u) && lane == 0) dst[n - 1u] = (OutT)lds_u16(stage_addr + 2u * (n - 1u));
    } else {
      if (lane == 0) dst[0] = (OutT)lds_u16(stage_addr);
      uint32_t *d32 = reinterpret_cast<uint32_t *>(dst + 1);
      const uint32_t nw = (n - 1u) >> 1;
#pragma unroll 4
      for (uint32_t w = lane; w < nw; w += 32u) {
        const uint32_t a = lds_u32(stage_addr + 4u * w), b = lds_u32(stage_addr + 4u * w + 4u);
        stg_cs_u32(d32 + w, __byte_perm(a, b, 0x5432));
      }
      if (((n - 1u) & 1u) && lane == 0) dst[n - 1u] = (OutT)lds_u16(stage_addr + 2u * (n - 1u));
    }
  }
}

// an all-ASCII interior tile: nothing was staged, the bytes come back from L2 and are widened on the way out
template <bool W32, bool BE, uint32_t TILE_BYTES, class OutT>
__device__ __forceinline__ void copy_out_ascii(const uint8_t *src, OutT *dst, unsigned lane) {
  if (W32) {
#pragma unroll 4
    for (uint32_t e = lane; e < TILE_BYTES; e += 32u) stg_cs_u32(dst + e, (uint32_t)__ldg(src + e));
  } else {
    constexpr uint32_t sel = BE ? 0x1404u : 0x4140u;  // bytes b0, b1 -> units
    if ((reinterpret_cast<uintptr_t>(dst) & 2u) == 0u) {
      uint32_t *d32 = reinterpret_cast<uint32_t *>(dst);
      const uint16_t *s16 = reinterpret_cast<const uint16_t *>(src);
#pragma unroll 4
      for (uint32_t w = lane; w < TILE_BYTES / 2u; w += 32u) stg_cs_u32(d32 + w, __byte_perm((uint32_t)__ldg(s16 + w), 0u, sel));
    } else {
      if (lane == 0) dst[0] = (OutT)(BE ? (uint32_t)__ldg(src) << 8 : (uint32_t)__ldg(src));
      uint32_t *d32 = reinterpret_cast<uint32_t *>(dst + 1);
#pragma unroll 4
      for (uint32_t w = lane; w < TILE_BYTES / 2u - 1u; w += 32u) {
        const uint32_t x = (uint32_t)__ldg(src + 2u * w + 1u) | ((uint32_t)__ldg(src + 2u * w + 2u) << 8);
        stg_cs_u32(d32 + w, __byte_perm(x, 0u, sel));
      }
      if (lane == 0) dst[TILE_BYTES - 1u] = (OutT)(BE ? (uint32_t)__ldg(src + TILE_BYTES - 1u) << 8 : (uint32_t)__ldg(src + TILE_BYTES - 1u));
    }
  }
}

// what a worker remembers of a tile whose copy-out is still due
struct PendingTile {
  uint32_t wtot = 0, tile = 0, iter = 0;
  bool valid = false, ascii = false;
};

template <int K, int MINB, bool W32, bool BE, int NW, bool DBG = false, int kAhead = 1>
__global__ void __launch_bounds__((NW + 1) * 32, MINB)
k_utf8_transcode_v3(const char *ptr, size_t len, typename std::conditional<W32, uint32_t, uint16_t>::type *out,
                    unsigned long long *desc, uint32_t epoch, uint32_t num_tiles, uint32_t num_cta_tiles, Scratch *scr,
                    ResultPOD *res, unsigned long long *dbg, unsigned long long *ts) {
  if (!DBG) dbg = nullptr;  // the clock64 instrumentation (tools/dbg_timing.py) exists in the DBG instantiation only
  using Gm = Geom3<K, W32, NW>;
  using OutT = typename std::conditional<W32, uint32_t, uint16_t>::type;
  constexpr uint32_t kUB = Gm::kUnitBytes;
  extern __shared__ __align__(16) uint32_t smem[];  // [NW][2] staging buffers
  __shared__ sp::Rings rg;  // hand-off rings, slot i & 3 for the CTA's i-th tile
  static_assert(NW <= 31, "one scan warp lane per worker");
  const InView in = make_view32(ptr, len);
  const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  const uint32_t mb = (uint32_t)__cvta_generic_to_shared(rg.mbar);
  constexpr uint32_t kMbTicket = sp::kMbTicket, kMbGoff = sp::kMbGoff, kMbTotals = sp::kMbTotals;
  if (threadIdx.x == 0) sp::init_rings(rg, NW);
  __syncthreads();

  if (warp == (unsigned)NW) {
    // ================================ scan warp ================================
    sp::scan_warp<NW, kAhead>(rg, desc, epoch, num_cta_tiles, scr, dbg, ts);
  } else {
    // ================================ workers ================================
    long long dbg_wt = 0, dbg_p1 = 0, dbg_wg = 0, dbg_copy = 0, dbg_p2 = 0, dbg_n = 0;
    const uint32_t stage0 = (uint32_t)__cvta_generic_to_shared(smem) + warp * 2u * Gm::kStageBytes;
    const bool poison = starts_with_continuation(in);
    const uint32_t one = (blockDim.x >> 5) - (uint32_t)NW;  // 1, but not a constant the assembler can fold (see bpd::bump)
    PendingTile q1, q2;  // tiles i - 1 and i - 2

    // waits for the offsets of a pending tile and copies it out of its staging buffer
    auto copy_out = [&](const PendingTile &q) {
      const uint32_t qs = q.iter & 3u;
      mbar_wait_hint(mb + kMbGoff + 8u * qs, (q.iter >> 2) & 1u);  // always: a worker never runs ahead of the offsets ring
      if (q.wtot) {
        OutT *dst = out + rg.goff[qs][warp];
        if (q.ascii)
          copy_out_ascii<W32, BE, Gm::kTileBytes>(reinterpret_cast<const uint8_t *>(in.base) + (unsigned long long)q.tile * Gm::kTileBytes, dst, lane);
        else
          copy_out_staged<W32>(stage0 + (q.iter & 1u) * Gm::kStageBytes, q.wtot, dst, lane);
      }
      __syncwarp();  // the staging buffer is about to be rewritten
    };

    for (uint32_t iter = 0;; iter++) {
      const uint32_t slot = iter & 3u, ph = (iter >> 2) & 1u;
      const long long w0 = dbg ? clock64() : 0;
      mbar_wait_hint(mb + kMbTicket + 8u * slot, ph);
      const long long w1 = dbg ? clock64() : 0;
      const uint32_t ct = rg.ticket[slot];
      if (ct >= num_cta_tiles) {  // CTA-uniform: drain
        if (q2.valid) copy_out(q2);
        if (q1.valid) copy_out(q1);
        break;
      }
      const uint32_t tile = ct * (uint32_t)NW + warp;
      const uint32_t stage_cur = stage0 + (iter & 1u) * Gm::kStageBytes;
      const bool active = tile < num_tiles;                                         // warp-uniform
      const unsigned long long t0 = (unsigned long long)tile * Gm::kTileBytes;     // virtual byte offsets from in.base
      const unsigned long long r0 = t0 + (unsigned long long)lane * Gm::kRegionBytes;
      const bool interior = active && t0 >= in.vbeg + 16ull && t0 + Gm::kTileBytes + 16ull <= in.vend;
      // ---- this lane's 32K contiguous bytes, the word before them and the byte after them ----
      uint32_t B[K][8];
      uint32_t pw = 0, nbyte = 0;
      if (interior) {
        const uint4 *gp = in.base + (r0 >> 4);
#pragma unroll
        for (int j = 0; j < K; j++) ldg_v8(gp + 2 * j, B[j]);  // one 256-bit load per block: half the LSU wavefronts of two 128-bit ones
        // the word before / the byte after the region sit in the neighbour lanes' registers: a strided load of them would
        // touch 16 lines per warp instruction (ncu: the three loads of this kernel were 130 of its 400 LSU wavefronts
        // per tile); only lanes 0 and 31 go to memory
        if (lane == 0) pw = __ldg(reinterpret_cast<const uint32_t *>(in.base) + (r0 >> 2) - 1);
        if (lane == 31) nbyte = __ldg(reinterpret_cast<const uint8_t *>(in.base) + r0 + Gm::kRegionBytes);
        const uint32_t up = __shfl_up_sync(kFull, B[K - 1][7], 1), dn = __shfl_down_sync(kFull, B[0][0], 1);
        if (lane != 0) pw = up;
        if (lane != 31) nbyte = dn & 0xFFu;
      } else if (active) {
#pragma unroll
        for (int j = 0; j < K; j++) {
          bool ins;
          load_granule(in, (r0 >> 4) + 2ull * j, &B[j][0], ins);
          load_granule(in, (r0 >> 4) + 2ull * j + 1ull, &B[j][4], ins);
        }
        pw = load_word_guarded(in, (long long)(r0 >> 2) - 1);
        const unsigned long long np = r0 + Gm::kRegionBytes;
        nbyte = (np >= in.vbeg && np < in.vend) ? (uint32_t)__ldg(reinterpret_cast<const uint8_t *>(in.base) + np) : 0u;
      } else {
#pragma unroll
        for (int j = 0; j < K; j++) {
#pragma unroll
          for (int i = 0; i < 8; i++) B[j][i] = 0u;
        }
      }
      // ---- pass 1: planes, emit masks, counts ----
      uint32_t hi = pw;
#pragma unroll
      for (int j = 0; j < K; j++) {
#pragma unroll
        for (int i = 0; i < 8; i++) hi |= B[j][i];
      }
      // the no-staging shortcut is for whole interior tiles only (an edge tile's bytes do not start at staging index 0)
      const bool ascii = interior && !poison && !__any_sync(kFull, (hi & kH) != 0u);
      uint32_t em[K];
      uint32_t cnt = 0;
      bp::Carry carry;
      if (!ascii) {
        carry = bp::carry_from_word(pw);
        uint32_t prev_l4 = carry.l4;
#pragma unroll
        for (int j = 0; j < K; j++) {
          const uint32_t nb = (j + 1 < K) ? B[(j + 1 < K) ? j + 1 : j][0] : nbyte;  // read BEFORE block j+1 is transposed
          const uint32_t next_nc = ((nb & 0xC0u) != 0x80u) ? 1u : 0u;
          bp::transpose_in(B[j]);
          uint32_t m = W32 ? bp::emit32_mask(B[j], next_nc) : bp::emit16_mask(B[j], prev_l4, next_nc);
          prev_l4 = B[j][7] & B[j][6] & B[j][5] & B[j][4];
          if (!interior) m &= active ? range_mask32(in, r0 + 32ull * j) : 0u;
          if (poison) m = 0;
          em[j] = m;
          cnt += (uint32_t)__popc(m);
        }
      } else {
#pragma unroll
        for (int j = 0; j < K; j++) em[j] = 0xFFFFFFFFu;
        cnt = Gm::kRegionBytes;
      }
      const uint32_t incl = bpd::warp_inclusive_u32(cnt);
      const uint32_t wtot = __shfl_sync(kFull, incl, 31);
      const uint32_t excl = incl - cnt;
      uint32_t tn = 0;
      bool took = false;
      if (lane == 0) {
        rg.tot[slot][warp] = wtot;
        // the worker that arrives last publishes the CTA aggregate (it never waits for the scan warp, header comment)
        // and reserves the CTA's next tile; the ticket's round trip hides behind its copy-out
        const uint32_t old = atomicAdd(&rg.acc[slot], (1u << 24) | wtot);
        if ((old >> 24) == (uint32_t)NW - 1u) {
          tn = atom_add_u32(&scr->ticket, 1u);
          took = true;
          rg.acc[slot] = 0;
          if (dbg) {
            unsigned long long now;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
            ts[ct] = now;
          }
          st_relaxed_u64(desc + ct, desc_pack(epoch, kStatusAggregate, 0, (old & 0xFFFFFFu) + wtot));
        }
        mbar_arrive(mb + kMbTotals + 8u * slot);
      }
      took = __any_sync(kFull, took);  // warp-uniform: this warp posts the ticket below
      auto post_ticket = [&]() {
        tn = __shfl_sync(kFull, tn, 0);
        if (lane == 0) {
          rg.ticket[(iter + (uint32_t)kAhead) & 3u] = tn;
          mbar_arrive(mb + kMbTicket + 8u * ((iter + (uint32_t)kAhead) & 3u));
        }
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
      const long long w2 = dbg ? clock64() : 0;
      // ---- the tile before the previous one leaves its staging buffer, which is this tile's ----
      long long w3 = w2;
      if (q2.valid) {
        if (dbg) {  // split the wait from the copy
          mbar_wait_hint(mb + kMbGoff + 8u * (q2.iter & 3u), (q2.iter >> 2) & 1u);
          w3 = clock64();
        }
        copy_out(q2);
      }
      if (took) post_ticket();
      const long long w4 = dbg ? clock64() : 0;
      // ---- pass 2: unit planes, transposition back, compaction into the staging buffer at alignment zero ----
      if (active && !ascii) {
        uint32_t badblocks = 0;
        uint32_t spa = stage_cur + kUB * excl;
#pragma unroll
        for (int j = 0; j < K; j++) {
          // four independent store chains (positions 0-7, 8-15, 16-23, 24-31): a chain's address register can
          // only advance once the store before it has read it, so one chain alone would serialise the block
          const uint32_t m = em[j];
          uint32_t s0 = spa;
          uint32_t s1 = spa + kUB * (uint32_t)__popc(m & 0xFFu);
          uint32_t s2 = spa + kUB * (uint32_t)__popc(m & 0xFFFFu);
          uint32_t s3 = spa + kUB * (uint32_t)__popc(m & 0xFFFFFFu);
          spa += kUB * (uint32_t)__popc(m);
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
            // One store chain per byte of the emit mask, its eight predicates extracted together: ptxas turns that into
            // one R2P + one LOP3 where a test per position costs eight LOP3 on the ALU pipe (measured: 0.885 -> 0.868 ms
            // per GiB).  Chain c covers positions 8c .. 8c + 7; its address advances on the FMA pipe.
#pragma unroll
            for (int c = 0; c < 4; c++) {
              bool pr[8];
#pragma unroll
              for (int i = 0; i < 8; i++) pr[i] = ((m >> (8 * c + i)) & 1u) != 0u;
              uint32_t sc = c == 0 ? s0 : (c == 1 ? s1 : (c == 2 ? s2 : s3));
#pragma unroll
              for (int i = 0; i < 8; i++) {
                if (pr[i]) {
                  sts_u16(sc, c < 2 ? U[8 * c + i] : __umulhi(U[8 * (c - 2) + i], 65536u));
                  sc = bpd::bump<2>(sc, one);
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
      }
      __syncwarp();  // the staged elements are visible to the whole warp
      if (dbg) {
        const long long w5 = clock64();
        dbg_wt += w1 - w0; dbg_p1 += w2 - w1; dbg_wg += w3 - w2; dbg_copy += w4 - w3; dbg_p2 += w5 - w4; dbg_n++;
      }
      q2 = q1;
      q1.valid = true; q1.wtot = wtot; q1.tile = tile; q1.iter = iter; q1.ascii = ascii;
    }
    if (dbg && lane == 0 && warp == 0) {
      unsigned long long *o = dbg + 16ull * blockIdx.x;
      o[8] = dbg_n; o[9] = dbg_wt; o[10] = dbg_p1; o[11] = dbg_wg; o[12] = dbg_copy; o[13] = dbg_p2;
    }
  }

  if (grid_last_thread(scr)) {
    const unsigned long long total = num_cta_tiles ? desc_value(ld_relaxed_u64(desc + (num_cta_tiles - 1u))) : 0ull;
    bpd::write_result_from_key(res, ld_relaxed_u64(&scr->err_key), total);
    scratch_reset(scr);
  }
}

inline size_t tiles_for(const void *in, size_t len_bytes, int k, unsigned align_mask = 15u) {
  const size_t span = (reinterpret_cast<uintptr_t>(in) & align_mask) + len_bytes;
  const size_t per_tile = (size_t)1024 * k;
  return (span + per_tile - 1) / per_tile;
}
inline size_t cta_tiles_for(size_t tiles) { return (tiles + kWorkers - 1) / kWorkers; }

template <int K, int MINB, bool W32, bool BE, int NW, bool DBG = false, int kAhead = 1>
cudaError_t launch_v3(const LaunchCtx &c, const char *in, size_t len, void *out, void *res) {
  using Gm = Geom3<K, W32, NW>;
  using OutT = typename std::conditional<W32, uint32_t, uint16_t>::type;
  static_assert(NW >= kWorkers, "the workspace is sized for CTA-tiles of at least kWorkers warp-tiles");
  const size_t tiles = tiles_for(in, len, K, 31u), cta_tiles = (tiles + NW - 1) / NW;
  if (cta_tiles + 1 > c.desc_capacity || tiles > 0xFFFFFF00ull) return cudaErrorInvalidValue;
  static KernelCache kc;
  int per_sm = 1;
  cudaError_t e = kernel_per_sm(kc, c.device, k_utf8_transcode_v3<K, MINB, W32, BE, NW, DBG, kAhead>, Gm::kThreads, Gm::kSmemBytes, &per_sm);
  if (e != cudaSuccess) return e;
  const size_t cap = (size_t)c.sm_count * per_sm;
  const unsigned grid = (unsigned)(cta_tiles < cap ? (cta_tiles ? cta_tiles : 1) : cap);
  k_utf8_transcode_v3<K, MINB, W32, BE, NW, DBG, kAhead><<<grid, Gm::kThreads, Gm::kSmemBytes, c.stream>>>(
      in, len, static_cast<OutT *>(out), c.desc, c.epoch, (uint32_t)tiles, (uint32_t)cta_tiles, c.scratch,
      static_cast<ResultPOD *>(res),
      reinterpret_cast<unsigned long long *>(((unsigned long long)(uint32_t)tuning(kTuneDbgHi) << 32) | (uint32_t)tuning(kTuneDbgLo)), c.cnt);
  count_launch(1);
  return cudaGetLastError();
}

}  // namespace

// Workspace, in 8-byte descriptor slots, the kernel needs for an input of `len` bytes: one descriptor per CTA-tile.
size_t utf8_to_utf16_tiles(const void *in, size_t len) { return cta_tiles_for(tiles_for(in, len, 1, 31u)) + 2; }
size_t utf8_to_utf32_tiles(const void *in, size_t len) { return cta_tiles_for(tiles_for(in, len, 1, 31u)) + 2; }

cudaError_t launch_convert_utf8_to_utf16(const LaunchCtx &c, const char *in, size_t len, uint16_t *out, void *res,
                                         bool big_endian) {
  // ONE CTA per SM: 16 worker warps (four per warp scheduler) + the scan warp, 96 bytes per lane (3 KiB warp-tiles, 48 KiB
  // CTA-tiles), 96 registers.  Measured on B200, 1 GiB of the mixed distribution, ms per launch (K = blocks per lane,
  // workers x CTAs per SM):
  //   K = 2:  7 x 3  0.934   11 x 2  0.868   12 x 2  0.893   13 x 2  0.954   16 x 1  0.917   20 x 1  0.856   24 x 1  0.844
  //   K = 3: 11 x 1  0.963   12 x 1  0.890   15 x 1  0.826   16 x 1  0.786   17 x 1  0.924   18 x 1  0.847
  //   K = 1: 15 x 2  1.021;  K = 4: 12 x 1  0.967
  // The workers of a CTA are coupled through the tile hand-offs, so the scheduler with the most workers sets the pace:
  // 16 (4-4-4-4) beats 15 and 17; larger tiles amortise the per-tile hand-offs, and the fatter warps (96-126 registers)
  // make up for the lower occupancy.
  if (big_endian) return launch_v3<3, 1, false, true, 16>(c, in, len, out, res);
  if (tuning(kTuneConvVariant) == 8) return launch_v3<3, 1, false, false, 16, true>(c, in, len, out, res);  // clock64 instrumentation (tools/dbg_timing.py)
  return launch_v3<3, 1, false, false, 16>(c, in, len, out, res);
}

cudaError_t launch_convert_utf8_to_utf32(const LaunchCtx &c, const char *in, size_t len, uint32_t *out, void *res) {
  // 32-bit elements double the staging buffers: 64 bytes per lane, 12 workers (three per scheduler), one CTA per SM
  // (1.052 ms per GiB; one block per lane, 11 x 2: 1.169; 16 x 1: 1.255; 24 x 1: 1.100)
  return launch_v3<2, 1, true, false, 12>(c, in, len, out, res);
}

}  // namespace b200
