// k_utf8_to_utf16.cu — sm_100a kernel K3: convert_utf8_to_utf16le/be[_with_errors], convert_utf8_to_utf32[_with_errors]
// (reference include/simdutf/implementation.h:3709-3800; semantics src/scalar/utf8_to_utf16/utf8_to_utf16.h:128-255,
// src/scalar/utf8_to_utf32/utf8_to_utf32.h:106-212).
//
// ONE launch, the input crosses HBM once (round 1 shipped a counts kernel + a transcoding kernel: the text was read
// twice, which capped the roofline fraction at 2/3).  A persistent grid hands out CTA-tiles (8 warps x 32 lanes x K
// blocks x 32 contiguous bytes = 16 KiB for K = 2) in increasing order through an atomic ticket.  Per CTA-tile:
//
//   pass 1   every lane loads its 32K contiguous bytes, transposes each 32-byte block into 8 BIT PLANES (bitplane.h)
//            and derives the block's emit mask; popcounts -> lane count -> warp inclusive scan -> warp total.
//   scan     the eight warp totals meet in shared memory; warp 0 publishes the CTA aggregate and resolves the tile's
//            global output offset with a decoupled look-back over epoch-tagged descriptors.  The look-back window is
//            128 descriptors per round (4 per lane): with ~450 CTAs in flight a tile's inclusive prefix appears
//            about one L2 round trip after its aggregate, i.e. 50-90 tiles behind the newest ticket, so a 32-wide
//            window needs several dependent rounds and never catches up (round 1's experiment stalled 56-86 % of
//            its samples there), while a 128-wide window finishes in one.
//   pass 2   bp::utf8_to_utf16_block (validation detector + the 16 planes of the candidate unit of all 32
//            positions, ~75 bitwise instructions per block), transposition back to 16-bit units, compaction with
//            predicated 16-bit shared stores into the WARP's contiguous staging buffer (lane l starts at the warp
//            prefix of the lane counts, shifted so that staging vectors line up with 16-byte-aligned output
//            addresses), then a warp-cooperative copy-out: LDS.128 -> STG.128, 512 contiguous bytes per
//            instruction; only the first and last partial vector of a warp-tile are written element-wise.
//
// The planes stay in registers across the scan, so counting costs nothing beyond what the transcoder needs anyway.
//
// Emission rule (bitplane.h): a unit is emitted at the LAST byte of its character (high surrogates at the third
// byte of a 4-byte sequence), so everything except one "is the next byte a continuation" bit looks backwards.
#include <type_traits>

#include "bitplane.h"
#include "bp_device.cuh"
#include "device_common.cuh"
#include "launch.h"

namespace b200 {

namespace {

using bpd::kThreads;
using bpd::kWarpsPerCta;
using bpd::sts_u16;
using bpd::sts_u32;

// W32 = false: UTF-16 output (16-bit units, 8 per 16-byte vector); W32 = true: UTF-32 (4 per vector).
template <int K, bool W32>
struct Geom {
  static constexpr uint32_t kRegionBytes = 32u * K;              // contiguous input bytes per lane
  static constexpr uint32_t kTileBytes = 32u * kRegionBytes;     // per warp
  static constexpr uint32_t kCtaTileBytes = 7u * kTileBytes;           // kWorkers warp-tiles
  static constexpr uint32_t kVec = W32 ? 4u : 8u;                // output elements per 16-byte vector
  static constexpr uint32_t kUnitBytes = W32 ? 4u : 2u;
  // a warp emits at most one element per input byte, in front of which sit up to kVec-1 elements of alignment pad
  static constexpr uint32_t kStageBytes = ((kTileBytes + kVec) * kUnitBytes + 15u) & ~15u;
  static constexpr uint32_t kSmemBytes = 7u * (kStageBytes + kTileBytes);  // per worker warp: staging + plane stash
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

__device__ __forceinline__ uint4 lds_v4(uint32_t addr) {
  uint4 r;
  asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(addr));
  return r;
}

// ---------------------------------------------------------------------------------------------
// Decoupled look-back over CTA-tile descriptors, 128 descriptors per round (lane l reads the four descriptors
// base - 4l .. base - 4l - 3).  Called by all 32 lanes of ONE warp; returns the exclusive prefix of `tile` (the number
// of elements every earlier tile emits) and publishes the tile's inclusive prefix.  Tiles are handed out in
// increasing order by an atomic ticket, so every predecessor is finished, running, or reserved by a CTA whose
// current tile is smaller still: the waits are bounded.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long cta_lookback128(unsigned long long *desc, uint32_t epoch, uint32_t tile,
                                                              uint32_t agg, Scratch *scr) {
  const unsigned lane = threadIdx.x & 31u;
  if (tile == 0) {
    if (lane == 0) st_relaxed_u64(desc, desc_pack(epoch, kStatusPrefix, 0, agg));
    return 0ull;
  }
  if (lane == 0) st_relaxed_u64(desc + tile, desc_pack(epoch, kStatusAggregate, 0, agg));
  unsigned long long sum = 0;
  long long base = (long long)tile - 1;
  while (true) {
    unsigned long long d[4];
    const long long i0 = base - 4ll * (long long)lane;
#pragma unroll
    for (int j = 0; j < 4; j++) d[j] = (i0 - j >= 0) ? ld_relaxed_u64(desc + (i0 - j)) : desc_pack(epoch, kStatusPrefix, 0, 0);
    for (uint32_t spins = 0;; spins++) {
      bool ready = true;
#pragma unroll
      for (int j = 0; j < 4; j++) ready = ready && desc_epoch(d[j]) == epoch && desc_status(d[j]) != 0u;
      if (ready) break;
      if (spins > (1u << 22)) {  // cannot happen (see above); never hang the device on a logic error
        report_error(scr, err_key(0, kOther));
        break;
      }
      __nanosleep(20);
#pragma unroll
      for (int j = 0; j < 4; j++)
        if (i0 - j >= 0) d[j] = ld_relaxed_u64(desc + (i0 - j));
    }
    // lane-local walk from the nearest descriptor backwards, up to and including the first inclusive prefix
    uint32_t aggs = 0;
    unsigned long long pref = 0;
    bool found = false;
#pragma unroll
    for (int j = 0; j < 4; j++) {
      if (!found) {
        if (desc_status(d[j]) == kStatusPrefix) {
          pref = desc_value(d[j]);
          found = true;
        } else {
          aggs += (uint32_t)desc_value(d[j]);
        }
      }
    }
    const unsigned pm = __ballot_sync(kFull, found);
    const unsigned first = pm ? (unsigned)(__ffs((int)pm) - 1) : 32u;
    sum += (unsigned long long)__reduce_add_sync(kFull, lane <= first ? aggs : 0u);
    if (pm) {
      const uint32_t lo = __shfl_sync(kFull, (uint32_t)pref, first), hi = __shfl_sync(kFull, (uint32_t)(pref >> 32), first);
      sum += ((unsigned long long)hi << 32) | lo;
      break;
    }
    base -= 128;
  }
  if (lane == 0) st_relaxed_u64(desc + tile, desc_pack(epoch, kStatusPrefix, 0, sum + agg));
  return sum;
}

// ---------------------------------------------------------------------------------------------
// K3: the single-pass bit-plane transcoder.
// BE (UTF-16 only): big-endian units — the low and high byte planes of the unit trade places before the
// transposition back (free), the ASCII paths put the byte into the upper half.
//
// Warp roles.  kWorkers = 7 worker warps transcode; warp 7 is the SCAN warp: it takes the tickets, publishes the CTA
// aggregates, runs the look-backs and hands the tile offsets to the workers.  The workers are software-pipelined by one
// tile: in iteration i they run pass 1 of tile i (planes, masks, counts), park the planes in a lane-private shared-
// memory stash, and then run pass 2 of tile i-1, whose offset the scan warp resolved while they were busy.  Nobody
// waits for a look-back: a chained scan makes tile t wait for the SLOWEST of its in-flight predecessors to publish,
// and with ~590 CTAs in flight that straggler costs several microseconds per tile (measured: 1.2-1.5 ms per GiB with
// the look-back on the workers' critical path, against ~0.9 for the transcoding itself); one tile of slack absorbs it.
// Producer/consumer hand-offs use mbarriers in shared memory, never a CTA-wide barrier.
// ---------------------------------------------------------------------------------------------
constexpr int kWorkers = 7;
constexpr int kScanWarp = kWorkers;
// Hand-offs go through mbarrier objects in shared memory (one arrival releases any number of waiters, and waiters do not
// wait for EACH OTHER the way the threads of a bar.sync do): a worker that is ahead never waits for a slower worker,
// only for the scan warp's data.
__device__ __forceinline__ void mbar_init(uint32_t addr, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(addr), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t addr) {  // release: this thread's earlier shared-memory writes are visible to waiters
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(addr) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t addr, uint32_t parity) {  // acquire
  uint32_t done;
  do {
    asm volatile(
        "{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
  } while (!done);
}
__device__ __forceinline__ void sts_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

template <int K, int MINB, bool W32, bool BE>
__global__ void __launch_bounds__(kThreads, MINB)
k_utf8_transcode_sp(const char *ptr, size_t len, typename std::conditional<W32, uint32_t, uint16_t>::type *out,
                    unsigned long long *desc, uint32_t epoch, uint32_t num_tiles, uint32_t num_cta_tiles, Scratch *scr,
                    ResultPOD *res) {
  using Gm = Geom<K, W32>;
  using OutT = typename std::conditional<W32, uint32_t, uint16_t>::type;
  constexpr uint32_t kUB = Gm::kUnitBytes;
  extern __shared__ __align__(16) uint32_t smem[];  // [kWorkers staging buffers][kWorkers plane stashes]
  __shared__ uint32_t s_tot[2][8];               // workers -> scan warp: the warp totals of tile i (slot i & 1)
  __shared__ unsigned long long s_goff[2][8];    // scan warp -> workers: every worker's global output offset for tile i
  __shared__ uint32_t s_ticket[2];               // scan warp -> workers: the CTA-tile of iteration i
  __shared__ __align__(8) unsigned long long s_mbar[6];  // [0,1] ticket posted, [2,3] offsets posted, [4,5] totals in
  const InView in = make_view16(ptr, len);
  const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  const uint32_t mb = (uint32_t)__cvta_generic_to_shared(s_mbar);
  if (threadIdx.x == 0) {
    mbar_init(mb + 0, 1); mbar_init(mb + 8, 1);
    mbar_init(mb + 16, 1); mbar_init(mb + 24, 1);
    mbar_init(mb + 32, kWorkers); mbar_init(mb + 40, kWorkers);
  }
  __syncthreads();

  if (warp == kScanWarp) {
    // ================================ scan warp ================================
    // Slot / phase discipline: the barrier of slot p = i & 1 completes its (i >> 1)-th phase for iteration i.  Nobody
    // can be two phases behind: the scan warp posts ticket i+2 only after every worker has delivered the totals of
    // tile i+1, i.e. has long consumed ticket i and offsets i-1.
    uint32_t t = 0;
    if (lane == 0) t = atomicAdd(&scr->ticket, 1u);
    t = __shfl_sync(kFull, t, 0);
    if (lane == 0) {
      s_ticket[0] = t;
      mbar_arrive(mb + 0);
    }
    for (uint32_t iter = 0; t < num_cta_tiles; iter++) {
      const uint32_t par = iter & 1u, ph = (iter >> 1) & 1u;
      mbar_wait(mb + 32 + 8 * par, ph);  // the workers' totals of tile t
      // reserve the next tile NOW: the workers pick it up a whole pass 2 later, so the ticket's round trip and the L2
      // prefetch of exactly that tile are off their critical path
      uint32_t tn = 0;
      if (lane == 0) tn = atomicAdd(&scr->ticket, 1u);
      tn = __shfl_sync(kFull, tn, 0);
      if (lane == 0) {
        s_ticket[par ^ 1u] = tn;
        mbar_arrive(mb + 8 * (par ^ 1u));
      }
      if (tn < num_cta_tiles) {
        const char *nx = reinterpret_cast<const char *>(in.base) + (unsigned long long)tn * Gm::kCtaTileBytes;
#pragma unroll
        for (uint32_t k = 0; k < (Gm::kCtaTileBytes + 4095u) / 4096u; k++) {
          const uint32_t off = k * 4096u + lane * 128u;
          if (off < Gm::kCtaTileBytes) asm volatile("prefetch.global.L2 [%0];" ::"l"(nx + off));
        }
      }
      const uint32_t mine = lane < (unsigned)kWorkers ? s_tot[par][lane] : 0u;
      const uint32_t incl = bpd::warp_inclusive_u32(mine);
      const uint32_t agg = __shfl_sync(kFull, incl, 31);
      const unsigned long long excl = cta_lookback128(desc, epoch, t, agg, scr);
      if (lane < (unsigned)kWorkers) s_goff[par][lane] = excl + (incl - mine);
      __syncwarp();
      if (lane == 0) mbar_arrive(mb + 16 + 8 * par);
      t = tn;
    }
  } else {
    // ================================ workers ================================
    const uint32_t stage_addr = (uint32_t)__cvta_generic_to_shared(smem) + warp * Gm::kStageBytes;
    const OutT *stage = reinterpret_cast<const OutT *>(smem) + (size_t)warp * (Gm::kStageBytes / kUB);
    const uint32_t stash_addr = (uint32_t)__cvta_generic_to_shared(smem) + kWorkers * Gm::kStageBytes + warp * Gm::kTileBytes + lane * 16u;
    const bool poison = starts_with_continuation(in);
    const unsigned long long out_units = (unsigned long long)(reinterpret_cast<uintptr_t>(out) / sizeof(OutT));
    const uint32_t one = blockDim.x >> 8;  // 1, but not a constant the assembler can fold (see bpd::bump)

    // the tile whose pass 2 is pending (per-lane: masks, lane offset, word before the region; warp-uniform: the rest)
    bool have_prev = false;
    uint32_t p_em[K], p_excl = 0, p_pw = 0, p_tile = 0, p_wtot = 0, p_par = 0, p_ph = 0;
    bool p_ascii = false, p_interior = false, p_active = false;
#pragma unroll
    for (int j = 0; j < K; j++) p_em[j] = 0;

    for (uint32_t iter = 0;; iter++) {
      const uint32_t par = iter & 1u, ph = (iter >> 1) & 1u;
      mbar_wait(mb + 8 * par, ph);
      const uint32_t ct = s_ticket[par];
      const bool more = ct < num_cta_tiles;  // CTA-uniform
      uint32_t B[K][8];
      uint32_t c_em[K], c_excl = 0, c_pw = 0, c_wtot = 0;
      bool c_ascii = false, c_interior = false, c_active = false;
      const uint32_t tile = ct * kWorkers + warp;
      if (more) {
        c_active = tile < num_tiles;                                               // warp-uniform
        const unsigned long long t0 = (unsigned long long)tile * Gm::kTileBytes;   // virtual byte offsets from in.base
        const unsigned long long r0 = t0 + (unsigned long long)lane * Gm::kRegionBytes;
        c_interior = c_active && t0 >= in.vbeg + 16ull && t0 + Gm::kTileBytes + 16ull <= in.vend;
        // ---- this lane's 32K contiguous bytes, the word before them and the byte after them ----
        uint32_t nbyte = 0;
        if (c_interior) {
          const uint4 *gp = in.base + (r0 >> 4);
#pragma unroll
          for (int j = 0; j < K; j++) {
            const uint4 v0 = __ldg(gp + 2 * j), v1 = __ldg(gp + 2 * j + 1);
            B[j][0] = v0.x; B[j][1] = v0.y; B[j][2] = v0.z; B[j][3] = v0.w;
            B[j][4] = v1.x; B[j][5] = v1.y; B[j][6] = v1.z; B[j][7] = v1.w;
          }
          c_pw = __ldg(reinterpret_cast<const uint32_t *>(in.base) + (r0 >> 2) - 1);
          nbyte = __ldg(reinterpret_cast<const uint8_t *>(in.base) + r0 + Gm::kRegionBytes);
        } else if (c_active) {
#pragma unroll
          for (int j = 0; j < K; j++) {
            bool ins;
            load_granule(in, (r0 >> 4) + 2ull * j, &B[j][0], ins);
            load_granule(in, (r0 >> 4) + 2ull * j + 1ull, &B[j][4], ins);
          }
          c_pw = load_word_guarded(in, (long long)(r0 >> 2) - 1);
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
        uint32_t hi = c_pw;
#pragma unroll
        for (int j = 0; j < K; j++) {
#pragma unroll
          for (int i = 0; i < 8; i++) hi |= B[j][i];
        }
        c_ascii = !__any_sync(kFull, (hi & kH) != 0u);
        uint32_t cnt = 0;
        if (!c_ascii) {
          uint32_t prev_l4;
          {
            uint32_t v[8];
            bp::planes_of_tail_word(c_pw, v);
            prev_l4 = v[7] & v[6] & v[5] & v[4];
          }
#pragma unroll
          for (int j = 0; j < K; j++) {
            const uint32_t nb = (j + 1 < K) ? B[(j + 1 < K) ? j + 1 : j][0] : nbyte;  // read BEFORE block j+1 is transposed
            const uint32_t next_nc = ((nb & 0xC0u) != 0x80u) ? 1u : 0u;
            bp::transpose_in(B[j]);
            uint32_t m = W32 ? bp::emit32_mask(B[j], next_nc) : bp::emit16_mask(B[j], prev_l4, next_nc);
            prev_l4 = B[j][7] & B[j][6] & B[j][5] & B[j][4];
            if (!c_interior) m &= range_mask32(in, r0 + 32ull * j);
            if (poison) m = 0;
            c_em[j] = m;
            cnt += (uint32_t)__popc(m);
          }
        } else {
#pragma unroll
          for (int j = 0; j < K; j++) {
            uint32_t m = c_active ? 0xFFFFFFFFu : 0u;
            if (!c_interior && c_active) m &= range_mask32(in, r0 + 32ull * j);
            if (poison) m = 0;
            c_em[j] = m;
            cnt += (uint32_t)__popc(m);
          }
        }
        const uint32_t incl = bpd::warp_inclusive_u32(cnt);
        c_wtot = __shfl_sync(kFull, incl, 31);
        c_excl = incl - cnt;
        if (lane == 0) {
          s_tot[par][warp] = c_wtot;
          mbar_arrive(mb + 32 + 8 * par);
        }
      } else {
#pragma unroll
        for (int j = 0; j < K; j++) c_em[j] = 0;
      }

      // ---- the planes of tile i go to the stash, those of tile i-1 come back ----
      if (more || have_prev) {
#pragma unroll
        for (int j = 0; j < K; j++) {
          uint4 v0 = make_uint4(0, 0, 0, 0), v1 = v0;
          if (have_prev) {
            v0 = lds_v4(stash_addr + (2u * j) * 512u);
            v1 = lds_v4(stash_addr + (2u * j + 1u) * 512u);
          }
          if (more) {
            sts_v4(stash_addr + (2u * j) * 512u, B[j][0], B[j][1], B[j][2], B[j][3]);
            sts_v4(stash_addr + (2u * j + 1u) * 512u, B[j][4], B[j][5], B[j][6], B[j][7]);
          }
          B[j][0] = v0.x; B[j][1] = v0.y; B[j][2] = v0.z; B[j][3] = v0.w;
          B[j][4] = v1.x; B[j][5] = v1.y; B[j][6] = v1.z; B[j][7] = v1.w;
        }
      }

      if (have_prev) {
        // ---- pass 2 of the pending tile ----
        mbar_wait(mb + 16 + 8 * p_par, p_ph);
        const unsigned long long goff = s_goff[p_par][warp];
        const unsigned long long r0 = (unsigned long long)p_tile * Gm::kTileBytes + (unsigned long long)lane * Gm::kRegionBytes;
        const uint32_t a_w = (uint32_t)((out_units + goff) & (Gm::kVec - 1u));  // offset of the tile inside a 16-byte output vector
        uint32_t badblocks = 0;
        bool staged = false;
        if (p_active) {
          // The running store address lives in a 32-bit shared-space register; it advances through the multiplier
          // (`one` * 2 + address) and the upper unit of a word is extracted with IMAD.HI, so that the compaction
          // costs the ALU pipe nothing but the predicate extraction.
          if (!p_ascii) {
            staged = true;
            bp::Carry carry = bp::carry_from_word(p_pw);
            uint32_t spa = stage_addr + kUB * (a_w + p_excl);
#pragma unroll
            for (int j = 0; j < K; j++) {
              // four independent store chains (positions 0-7, 8-15, 16-23, 24-31): a chain's address register can
              // only advance once the store before it has read it, so one chain alone would serialise the block
              const uint32_t m = p_em[j];
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
            }
          } else if (p_interior && !poison && a_w == 0u) {
            // every lane emits exactly 32K elements and the tile's output is vector-aligned: widen in registers and
            // store straight to global memory (nothing staged, no errors possible in an all-ASCII interior tile)
            uint4 *gv = reinterpret_cast<uint4 *>(out + goff + p_excl);
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
          } else {
            staged = true;
            uint32_t spa = stage_addr + kUB * (a_w + p_excl);
#pragma unroll
            for (int j = 0; j < K; j++) {
              const uint32_t m = p_em[j];
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
          if (!p_interior) {
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
        __syncwarp();
        // ---- staging -> global: the warp's elements [a_w, a_w + wtot) of the staging buffer go to out[goff ...] ----
        if (staged && p_wtot) {
          OutT *gbase = out + goff - a_w;  // 16-byte aligned
          const uint32_t end = a_w + p_wtot;
          const uint32_t v0 = a_w ? 1u : 0u, v1 = end / Gm::kVec;
          const uint32_t head_end = a_w ? (end < Gm::kVec ? end : Gm::kVec) : 0u;
          for (uint32_t v = v0 + lane; v < v1; v += 32u)
            stg_stream_v4(reinterpret_cast<uint4 *>(gbase) + v, lds_v4(stage_addr + 16u * v));
          if (lane >= a_w && lane < head_end) gbase[lane] = stage[lane];  // first partial vector (shared with the previous tile)
          const uint32_t ti = v1 * Gm::kVec + lane;                        // last partial vector (shared with the next tile)
          if (lane < Gm::kVec && ti >= head_end && ti < end) gbase[ti] = stage[ti];
        }
        __syncwarp();  // the staging buffer is rewritten by the next tile
      }
      if (!more) break;
      have_prev = true;
#pragma unroll
      for (int j = 0; j < K; j++) p_em[j] = c_em[j];
      p_excl = c_excl; p_pw = c_pw; p_tile = tile; p_wtot = c_wtot; p_par = par; p_ph = ph;
      p_ascii = c_ascii; p_interior = c_interior; p_active = c_active;
    }
  }

  if (grid_last_thread(scr)) {
    const unsigned long long total = num_cta_tiles ? desc_value(ld_relaxed_u64(desc + (num_cta_tiles - 1u))) : 0ull;
    bpd::write_result_from_key(res, ld_relaxed_u64(&scr->err_key), total);
    scratch_reset(scr);
  }
}


// ---------------------------------------------------------------------------------------------
// K3 v3: the same single pass with the look-back OFF the workers' path.
//
// In k_utf8_transcode_sp a worker needs the tile's global output offset BEFORE it compacts (the staging position carries
// the alignment of the output address), so the planes are parked in shared memory for one tile and every hiccup of the
// chained scan shows up as a wait in front of pass 2 (ncu: 24 % of all samples on that one mbarrier).  Here a worker
// transcodes and compacts tile i into staging buffer i & 1 at alignment ZERO, knowing nothing but the warp's own lane
// prefix; the global offset is needed only by the copy-out, and the copy-out of tile i runs TWO tiles later, between
// pass 1 and pass 2 of tile i + 2 (just before that buffer is overwritten): the look-back of a tile has two whole tile
// times to finish.  The copy-out realigns: the destination is only 2-byte aligned, so it goes out as 32-bit words (128
// contiguous bytes per warp instruction), one byte permute per word when the destination starts on an odd unit.
//
// Who does what (measured on B200 with clock64 instrumentation, 1 GiB of mixed text, 444 CTAs):
//   * the worker that delivers the LAST warp total of a tile publishes the CTA aggregate itself.  With the scan warp
//     publishing it, an aggregate waited for the scan warp's previous look-back, which waited for other CTAs'
//     aggregates, ...: 13 us and 25 polls per look-back; published by the workers, 5 us and 5 polls;
//   * the worker that reaches pass 2 FIRST reserves the CTA's next tile (global atomic ticket); the ticket's round trip
//     hides behind its pass 2.  Tickets keep the scan deadlock-free when not all CTAs are resident;
//   * the scan warp only looks back (coalesced: a warp load reads 32 consecutive descriptors) and posts the offsets.
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

__device__ __forceinline__ void mbar_wait_hint(uint32_t addr, uint32_t parity) {  // acquire; the hardware parks the thread
  uint32_t done;
  do {
    asm volatile(
        "{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3; selp.u32 %0, 1, 0, p; }"
        : "=r"(done)
        : "r"(addr), "r"(parity), "r"(1000000u)
        : "memory");
  } while (!done);
}
__device__ __forceinline__ uint32_t lds_u32(uint32_t addr) {
  uint32_t r;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(r) : "r"(addr));
  return r;
}
__device__ __forceinline__ uint32_t lds_u16(uint32_t addr) {
  uint32_t r;
  asm volatile("{ .reg .b16 t; ld.shared.b16 t, [%1]; cvt.u32.u16 %0, t; }" : "=r"(r) : "r"(addr));
  return r;
}
__device__ __forceinline__ void stg_cs_u32(void *p, uint32_t v) {
  asm volatile("st.global.cs.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t atom_add_u32(unsigned int *p, uint32_t v) {
  uint32_t r;
  asm volatile("atom.global.add.u32 %0, [%1], %2;" : "=r"(r) : "l"(p), "r"(v) : "memory");
  return r;
}

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

template <int K, int MINB, bool W32, bool BE, int NW, bool DBG = false>
__global__ void __launch_bounds__((NW + 1) * 32, MINB)
k_utf8_transcode_v3(const char *ptr, size_t len, typename std::conditional<W32, uint32_t, uint16_t>::type *out,
                    unsigned long long *desc, uint32_t epoch, uint32_t num_tiles, uint32_t num_cta_tiles, Scratch *scr,
                    ResultPOD *res, unsigned long long *dbg, unsigned long long *ts) {
  if (!DBG) dbg = nullptr;  // the clock64 instrumentation (tools/dbg_timing.py) exists in the DBG instantiation only
  using Gm = Geom3<K, W32, NW>;
  using OutT = typename std::conditional<W32, uint32_t, uint16_t>::type;
  constexpr uint32_t kUB = Gm::kUnitBytes;
  extern __shared__ __align__(16) uint32_t smem[];  // [NW][2] staging buffers
  // hand-off rings, slot i & 3 for the CTA's i-th tile
  __shared__ uint32_t s_tot[4][16];               // workers -> scan warp: the warp totals
  __shared__ unsigned long long s_goff[4][16];    // scan warp -> workers: every worker's global output offset
  __shared__ uint32_t s_ticket[4];               // the CTA-tile index
  __shared__ uint32_t s_acc[4];                  // arrivals << 16 | sum of the warp totals
  __shared__ uint32_t s_elect[4];                // workers that have reached pass 2
  __shared__ __align__(8) unsigned long long s_mbar[12];  // [0,4) ticket posted, [4,8) offsets posted, [8,12) totals in
  static_assert(NW <= 15, "one scan warp lane per worker; 16-bit packing of the warp totals");
  const InView in = make_view16(ptr, len);
  const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  const uint32_t mb = (uint32_t)__cvta_generic_to_shared(s_mbar);
  constexpr uint32_t kMbTicket = 0u, kMbGoff = 32u, kMbTotals = 64u;
  if (threadIdx.x == 0) {
#pragma unroll
    for (uint32_t k = 0; k < 4; k++) {
      mbar_init(mb + kMbTicket + 8u * k, 1);
      mbar_init(mb + kMbGoff + 8u * k, 1);
      mbar_init(mb + kMbTotals + 8u * k, NW);
      s_acc[k] = 0;
      s_elect[k] = 0;
    }
  }
  __syncthreads();

  if (warp == (unsigned)NW) {
    // ================================ scan warp ================================
    uint32_t t = 0;
    if (lane == 0) t = atomicAdd(&scr->ticket, 1u);
    t = __shfl_sync(kFull, t, 0);
    if (lane == 0) {
      s_ticket[0] = t;
      mbar_arrive(mb + kMbTicket);
    }
    long long dbg_wait = 0, dbg_lb = 0, dbg_lbmax = 0, dbg_polls = 0, dbg_n = 0, dbg_late = 0, dbg_seen = 0, dbg_start = 0;
    for (uint32_t iter = 0;; iter++) {
      const uint32_t slot = iter & 3u, ph = (iter >> 2) & 1u;
      if (iter) {  // the tickets after the first are taken by the workers
        mbar_wait_hint(mb + kMbTicket + 8u * slot, ph);
        t = s_ticket[slot];
      }
      if (t >= num_cta_tiles) break;
      const long long c0 = dbg ? clock64() : 0;
      mbar_wait_hint(mb + kMbTotals + 8u * slot, ph);
      const long long c1 = dbg ? clock64() : 0;
      unsigned long long now0 = 0;
      if (dbg) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now0));
      const uint32_t mine = lane < (unsigned)NW ? s_tot[slot][lane] : 0u;
      const uint32_t incl = bpd::warp_inclusive_u32(mine);
      const uint32_t agg = __shfl_sync(kFull, incl, 31);
      unsigned long long sum = 0;
      if (t == 0) {
        if (lane == 0) st_relaxed_u64(desc, desc_pack(epoch, kStatusPrefix, 0, agg));
      } else {
        // Look-back, COALESCED: descriptor base - 32 j - lane goes to lane `lane` of load j, so a warp load touches 256
        // contiguous bytes (8 sectors), and a lane polls only a descriptor that is not ready yet.  (A persistent grid
        // of equal tiles drifts into lockstep: then no predecessor of the current wave has its prefix yet and every
        // CTA reads the whole in-flight window, G descriptors G times per wave, on a few dozen L2 lines.)
        constexpr int kR = 4;
        long long base = (long long)t - 1;
        unsigned long long d[kR];
#pragma unroll
        for (int j = 0; j < kR; j++) {
          const long long idx = base - 32ll * j - (long long)lane;
          d[j] = idx >= 0 ? ld_relaxed_u64(desc + idx) : desc_pack(epoch, kStatusPrefix, 0, 0);
        }
        bool done = false;
        while (!done) {
#pragma unroll
          for (int j = 0; j < kR; j++) {
            if (!done) {  // warp-uniform
              const long long idx = base - 32ll * j - (long long)lane;
              uint32_t spins = 0;
              while (__any_sync(kFull, desc_epoch(d[j]) != epoch || desc_status(d[j]) == 0u)) {
                if (desc_epoch(d[j]) != epoch || desc_status(d[j]) == 0u) d[j] = ld_relaxed_u64(desc + idx);  // idx >= 0 here
                dbg_polls++;
                if (++spins > (1u << 24)) {  // cannot happen (tickets are handed out in order); never hang the device on a logic error
                  report_error(scr, err_key(0, kOther));
                  break;
                }
              }
              if (dbg && j == 0 && base == (long long)t - 1 && t >= 64) {
                unsigned long long now;
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
                const long long mine_ts = (long long)ld_relaxed_u64(ts + t), pred_ts = (long long)ld_relaxed_u64(ts + idx);
                long long late = pred_ts - mine_ts;  // > 0: this predecessor published after me
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                  const long long x = __shfl_xor_sync(kFull, late, o);
                  late = x > late ? x : late;
                }
                dbg_late += late; dbg_seen += (long long)now - (mine_ts + (late > 0 ? late : 0));
                dbg_start += (long long)now0 - mine_ts;
              }
              const unsigned pm = __ballot_sync(kFull, desc_status(d[j]) == kStatusPrefix);
              const unsigned first = pm ? (unsigned)(__ffs((int)pm) - 1) : 32u;
              sum += (unsigned long long)__reduce_add_sync(kFull, lane < first ? (uint32_t)desc_value(d[j]) : 0u);
              if (pm) {
                const uint32_t lo = __shfl_sync(kFull, (uint32_t)desc_value(d[j]), first);
                const uint32_t hi = __shfl_sync(kFull, (uint32_t)(desc_value(d[j]) >> 32), first);
                sum += ((unsigned long long)hi << 32) | lo;
                done = true;
              }
            }
          }
          if (!done) {
            base -= 32 * kR;
#pragma unroll
            for (int j = 0; j < kR; j++) {
              const long long idx = base - 32ll * j - (long long)lane;
              d[j] = idx >= 0 ? ld_relaxed_u64(desc + idx) : desc_pack(epoch, kStatusPrefix, 0, 0);
            }
          }
        }
        if (lane == 0) st_relaxed_u64(desc + t, desc_pack(epoch, kStatusPrefix, 0, sum + agg));
      }
      if (lane < (unsigned)NW) s_goff[slot][lane] = sum + (incl - mine);
      __syncwarp();
      if (lane == 0) mbar_arrive(mb + kMbGoff + 8u * slot);
      if (dbg) {
        const long long c2 = clock64();
        dbg_wait += c1 - c0; dbg_lb += c2 - c1; dbg_lbmax = (c2 - c1) > dbg_lbmax ? (c2 - c1) : dbg_lbmax; dbg_n++;
      }
    }
    if (dbg && lane == 0) {
      unsigned long long *o = dbg + 16ull * blockIdx.x;
      o[0] = dbg_n; o[1] = dbg_wait; o[2] = dbg_lb; o[3] = dbg_lbmax; o[4] = dbg_polls; o[5] = dbg_late; o[6] = dbg_seen; o[7] = dbg_start;
    }
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
        OutT *dst = out + s_goff[qs][warp];
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
      const uint32_t ct = s_ticket[slot];
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
        for (int j = 0; j < K; j++) {
          const uint4 v0 = __ldg(gp + 2 * j), v1 = __ldg(gp + 2 * j + 1);
          B[j][0] = v0.x; B[j][1] = v0.y; B[j][2] = v0.z; B[j][3] = v0.w;
          B[j][4] = v1.x; B[j][5] = v1.y; B[j][6] = v1.z; B[j][7] = v1.w;
        }
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
        s_tot[slot][warp] = wtot;
        // the worker that arrives last publishes the CTA aggregate (it never waits for the scan warp, header comment)
        // and reserves the CTA's next tile; the ticket's round trip hides behind its copy-out
        const uint32_t old = atomicAdd(&s_acc[slot], (1u << 16) | wtot);
        if ((old >> 16) == (uint32_t)NW - 1u) {
          tn = atom_add_u32(&scr->ticket, 1u);
          took = true;
          s_acc[slot] = 0;
          if (dbg) {
            unsigned long long now;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
            ts[ct] = now;
          }
          st_relaxed_u64(desc + ct, desc_pack(epoch, kStatusAggregate, 0, (old & 0xFFFFu) + wtot));
        }
        mbar_arrive(mb + kMbTotals + 8u * slot);
      }
      took = __any_sync(kFull, took);  // warp-uniform: this warp posts the ticket below
      auto post_ticket = [&]() {
        tn = __shfl_sync(kFull, tn, 0);
        if (lane == 0) {
          s_ticket[(iter + 1u) & 3u] = tn;
          mbar_arrive(mb + kMbTicket + 8u * ((iter + 1u) & 3u));
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

inline size_t tiles_for(const void *in, size_t len_bytes, int k) {
  const size_t span = (reinterpret_cast<uintptr_t>(in) & 15u) + len_bytes;
  const size_t per_tile = (size_t)1024 * k;
  return (span + per_tile - 1) / per_tile;
}
inline size_t cta_tiles_for(size_t tiles) { return (tiles + kWorkers - 1) / kWorkers; }

template <int K, int MINB, bool W32, bool BE>
cudaError_t launch_sp(const LaunchCtx &c, const char *in, size_t len, void *out, void *res) {
  using Gm = Geom<K, W32>;
  using OutT = typename std::conditional<W32, uint32_t, uint16_t>::type;
  const size_t tiles = tiles_for(in, len, K), cta_tiles = cta_tiles_for(tiles);
  if (cta_tiles + 1 > c.desc_capacity || tiles > 0xFFFFFF00ull) return cudaErrorInvalidValue;
  static KernelCache kc;
  int per_sm = 1;
  cudaError_t e = kernel_per_sm(kc, c.device, k_utf8_transcode_sp<K, MINB, W32, BE>, kThreads, Gm::kSmemBytes, &per_sm);
  if (e != cudaSuccess) return e;
  const size_t cap = (size_t)c.sm_count * per_sm;
  const unsigned grid = (unsigned)(cta_tiles < cap ? (cta_tiles ? cta_tiles : 1) : cap);
  k_utf8_transcode_sp<K, MINB, W32, BE><<<grid, kThreads, Gm::kSmemBytes, c.stream>>>(
      in, len, static_cast<OutT *>(out), c.desc, c.epoch, (uint32_t)tiles, (uint32_t)cta_tiles, c.scratch,
      static_cast<ResultPOD *>(res));
  count_launch(1);
  return cudaGetLastError();
}

template <int K, int MINB, bool W32, bool BE, int NW, bool DBG = false>
cudaError_t launch_v3(const LaunchCtx &c, const char *in, size_t len, void *out, void *res) {
  using Gm = Geom3<K, W32, NW>;
  using OutT = typename std::conditional<W32, uint32_t, uint16_t>::type;
  static_assert(NW >= kWorkers, "the workspace is sized for CTA-tiles of at least kWorkers warp-tiles");
  const size_t tiles = tiles_for(in, len, K), cta_tiles = (tiles + NW - 1) / NW;
  if (cta_tiles + 1 > c.desc_capacity || tiles > 0xFFFFFF00ull) return cudaErrorInvalidValue;
  static KernelCache kc;
  int per_sm = 1;
  cudaError_t e = kernel_per_sm(kc, c.device, k_utf8_transcode_v3<K, MINB, W32, BE, NW, DBG>, Gm::kThreads, Gm::kSmemBytes, &per_sm);
  if (e != cudaSuccess) return e;
  const size_t cap = (size_t)c.sm_count * per_sm;
  const unsigned grid = (unsigned)(cta_tiles < cap ? (cta_tiles ? cta_tiles : 1) : cap);
  k_utf8_transcode_v3<K, MINB, W32, BE, NW, DBG><<<grid, Gm::kThreads, Gm::kSmemBytes, c.stream>>>(
      in, len, static_cast<OutT *>(out), c.desc, c.epoch, (uint32_t)tiles, (uint32_t)cta_tiles, c.scratch,
      static_cast<ResultPOD *>(res),
      reinterpret_cast<unsigned long long *>(((unsigned long long)(uint32_t)tuning(kTuneDbgHi) << 32) | (uint32_t)tuning(kTuneDbgLo)), c.cnt);
  count_launch(1);
  return cudaGetLastError();
}

}  // namespace

// Workspace, in 8-byte descriptor slots, the kernel needs for an input of `len` bytes: one descriptor per CTA-tile.
size_t utf8_to_utf16_tiles(const void *in, size_t len) { return cta_tiles_for(tiles_for(in, len, 2)) + 2; }
size_t utf8_to_utf32_tiles(const void *in, size_t len) { return cta_tiles_for(tiles_for(in, len, 2)) + 2; }

cudaError_t launch_convert_utf8_to_utf16(const LaunchCtx &c, const char *in, size_t len, uint16_t *out, void *res,
                                         bool big_endian) {
  if (big_endian) return launch_v3<2, 3, false, true, 7>(c, in, len, out, res);
  switch (tuning(kTuneConvVariant)) {
    case 1: return launch_sp<2, 3, false, false>(c, in, len, out, res);
    case 4: return launch_v3<2, 2, false, false, 11>(c, in, len, out, res);
    case 8: return launch_v3<2, 3, false, false, 7, true>(c, in, len, out, res);
    case 9: return launch_v3<2, 2, false, false, 11, true>(c, in, len, out, res);
    default: return launch_v3<2, 3, false, false, 7>(c, in, len, out, res);
  }
}

cudaError_t launch_convert_utf8_to_utf32(const LaunchCtx &c, const char *in, size_t len, uint32_t *out, void *res) {
  if (tuning(kTuneConvVariant) == 1) return launch_sp<2, 2, true, false>(c, in, len, out, res);
  return launch_v3<2, 2, true, false, 7>(c, in, len, out, res);
}

}  // namespace b200
