// k_utf8_to_utf16.cu — sm_100a kernel K3: convert_utf8_to_utf16le[_with_errors]
// (reference include/simdutf/implementation.h:3709-3745; semantics src/scalar/utf8_to_utf16/utf8_to_utf16.h:128-255).
//
// One pass over the input, warp-specialised CTAs of 8 transcoding warps + 1 scan warp:
//
//   transcoding warp (owns 32*G granules = 512*G contiguous input bytes of the CTA's tile)
//     1. G coalesced 128-bit streaming loads per lane; the 3-byte look-behind / look-ahead comes from the
//        neighbouring lane by shuffle.
//     2. branch-free SWAR transcoder + validation detector (swar.h: u8_to_utf16_word), results kept in
//        registers (two candidate units per 32-bit register, one packed emit mask per granule).
//     3. popcount of the emit masks -> warp-level exclusive scan -> warp total posted to shared memory,
//        bar.arrive on barrier A  (the warp does NOT wait).
//     4. candidates compacted into the warp's own staging region of shared memory (warp-relative offsets:
//        no other warp's count is needed).
//     5. bar.sync on barrier B of the PREVIOUS tile: the scan warp has published where that tile's units
//        start in the output (it had a whole tile's worth of transcoding time to find out).
//     6. previous tile's staging region -> global memory as 16-byte vector stores, re-aligned on the fly
//        (two aligned 128-bit shared loads + a warp-uniform funnel shift), so the store alignment never has
//        to be known while the units are produced.  Staging regions are double-buffered.
//   scan warp
//     bar.sync A -> sums the 8 warp totals -> publishes the tile aggregate and runs the decoupled
//     look-back over the tile descriptors (128 predecessors per round) -> posts per-warp output offsets and
//     the ticket of the tile after next -> bar.arrive B.  Its latency (L2 round trips) overlaps the
//     transcoding of the next tile; A and B alternate between two barrier ids by tile parity.
//
// Tiles are handed out by an atomic ticket in increasing order (forward progress of the look-back).
// Emit rule and values: swar.h (one unit per non-continuation byte + one for the byte after a byte >= 0xF0,
// so an output buffer of utf16_length_from_utf8() units is never overrun, even for invalid input).
#include <cstdlib>

#include "device_common.cuh"
#include "launch.h"

namespace b200 {

namespace {

constexpr int kCW = 8;                       // transcoding warps per CTA
constexpr int kT16Threads = (kCW + 1) * 32;  // + the scan warp
constexpr int kBarA = 1, kBarB = 3;          // named barriers kBarA+parity, kBarB+parity (0 is __syncthreads)

// barrier ids are immediates (a register id makes ptxas reserve all 16 barriers): id = BASE + parity
template <int BASE>
__device__ __forceinline__ void bar_sync(uint32_t par, int n) {
  if (par) asm volatile("bar.sync %0, %1;" ::"n"(BASE + 1), "r"(n) : "memory");
  else asm volatile("bar.sync %0, %1;" ::"n"(BASE), "r"(n) : "memory");
}
template <int BASE>
__device__ __forceinline__ void bar_arrive(uint32_t par, int n) {
  if (par) asm volatile("bar.arrive %0, %1;" ::"n"(BASE + 1), "r"(n) : "memory");
  else asm volatile("bar.arrive %0, %1;" ::"n"(BASE), "r"(n) : "memory");
}

template <int G>
struct T16Smem {
  static constexpr uint32_t kWarpUnits = 32u * G * 16u;  // a warp emits at most one unit per input byte
  static constexpr uint32_t kRegion = kWarpUnits + 16u;  // 8 units of padding in front, 8 behind (vector over-read)
  alignas(16) uint16_t stage[2][kCW][kRegion];  // [tile parity]
  unsigned long long warp_goff[2][kCW];         // index in the output of each warp's first unit
  uint32_t warp_total[2][kCW];
  uint32_t tile_pub[2];                         // ticket of the tile after next, published with B
  uint32_t first_tiles[2];
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

// Warp-relative exclusive offsets of per-lane per-item counts c[j] <= 16, in element order (item, lane).
template <int G>
__device__ __forceinline__ uint32_t warp_exclusive_offsets(const uint32_t (&c)[G], uint32_t (&off)[G]) {
  const unsigned lane = threadIdx.x & 31u;
  uint32_t run = 0;
#pragma unroll
  for (int j = 0; j < G; j += 2) {
    const uint32_t c1 = (j + 1 < G) ? c[j + 1] : 0u;
    uint32_t incl = c[j] | (c1 << 16);  // two 16-bit lanes; 32 * 16 < 65536
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t t = __shfl_up_sync(kFull, incl, o);
      if (lane >= (unsigned)o) incl += t;
    }
    const uint32_t tot = __shfl_sync(kFull, incl, 31);
    off[j] = run + (incl & 0xFFFFu) - c[j];
    run += tot & 0xFFFFu;
    if (j + 1 < G) {
      off[j + 1] = run + (incl >> 16) - c1;
      run += tot >> 16;
    }
  }
  return run;
}

// Words q = 0..3 of the 8-unit window that starts S units into the 16 units {A, B}.
template <int S>
__device__ __forceinline__ uint4 funnel_units(const uint4 &A, const uint4 &B) {
  const uint32_t c[8] = {A.x, A.y, A.z, A.w, B.x, B.y, B.z, B.w};
  uint4 r;
  if (S % 2 == 0) {
    r.x = c[S / 2 + 0]; r.y = c[S / 2 + 1]; r.z = c[S / 2 + 2]; r.w = c[S / 2 + 3];
  } else {
    r.x = __byte_perm(c[S / 2 + 0], c[S / 2 + 1], 0x5432);
    r.y = __byte_perm(c[S / 2 + 1], c[S / 2 + 2], 0x5432);
    r.z = __byte_perm(c[S / 2 + 2], c[S / 2 + 3], 0x5432);
    r.w = __byte_perm(c[S / 2 + 3], c[(S / 2 + 4) & 7], 0x5432);
  }
  return r;
}

// Copies the warp's n staged units (stage[8 .. 8+n)) to dst[0 .. n).  Vector v of the 16-byte-aligned
// destination holds units 8v-a .. 8v-a+7 of the warp (a = units between the aligned base and dst).
template <int S>  // S = (8 - a) & 7, warp-uniform
__device__ __forceinline__ void copy_out_warp(const uint16_t *stage, uint16_t *dst, uint32_t n) {
  const unsigned lane = threadIdx.x & 31u;
  constexpr uint32_t a = (8u - S) & 7u;
  const uint32_t nvec = (a + n + 7u) >> 3;
  uint16_t *dbase = dst - a;  // 16-byte aligned
  const uint4 *sv = reinterpret_cast<const uint4 *>(stage);
  for (uint32_t v = lane; v < nvec; v += 32u) {
    const bool full = (v > 0 || a == 0) && (8u * v + 8u <= a + n);
    if (full) {
      uint4 r;
      if (S == 0) {
        r = sv[v + 1];
      } else {
        const uint4 A = sv[v], B = sv[v + 1];
        r = funnel_units<S>(A, B);
      }
      stg_stream_v4(reinterpret_cast<uint4 *>(dbase + 8u * v), r);
    } else {
#pragma unroll
      for (uint32_t t = 0; t < 8; t++) {
        const uint32_t e = 8u * v + t;  // element of the aligned destination
        if (e >= a && e < a + n) dbase[e] = stage[8u + e - a];
      }
    }
  }
}

// Copy-out of one staged warp-tile: n units at stage[8..8+n) -> out[goff .. goff+n).
__device__ __forceinline__ void flush_warp(const uint16_t *stage, uint16_t *out, unsigned long long goff, uint32_t n) {
  uint16_t *dst = out + goff;
  const uint32_t a = (uint32_t)(reinterpret_cast<uintptr_t>(dst) >> 1) & 7u;
  switch ((8u - a) & 7u) {
    case 0: copy_out_warp<0>(stage, dst, n); break;
    case 1: copy_out_warp<1>(stage, dst, n); break;
    case 2: copy_out_warp<2>(stage, dst, n); break;
    case 3: copy_out_warp<3>(stage, dst, n); break;
    case 4: copy_out_warp<4>(stage, dst, n); break;
    case 5: copy_out_warp<5>(stage, dst, n); break;
    case 6: copy_out_warp<6>(stage, dst, n); break;
    default: copy_out_warp<7>(stage, dst, n); break;
  }
}

// Steps 1-4 of one tile for one transcoding warp; returns the number of units the warp staged.
// EDGE: the tile touches the first or last byte of the buffer.
template <int G, bool EDGE>
__device__ __forceinline__ uint32_t transcode_tile(const InView &in, Scratch *scr, uint32_t tile, uint32_t par,
                                                   T16Smem<G> &sm) {
  const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  const unsigned long long g0 = ((unsigned long long)tile * kCW + warp) * (32ull * G);

  // ---- 1. load + neighbours ----
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

  // ---- 2. transcode into registers ----
  uint32_t U[G][8], M[G], cnt[G];
  uint32_t flagged = 0;
#pragma unroll
  for (int j = 0; j < G; j++) {
    const uint32_t hi = (w[j][0] | w[j][1] | w[j][2] | w[j][3] | pw[j]) & kH;
    if (!__any_sync(kFull, hi != 0u)) {
      // 512 ASCII bytes: units are the bytes, every position emits
#pragma unroll
      for (int k = 0; k < 4; k++) {
        U[j][2 * k] = __byte_perm(w[j][k], 0u, 0x4140);
        U[j][2 * k + 1] = __byte_perm(w[j][k], 0u, 0x4342);
      }
      M[j] = 0xF0F0F0F0u;
    } else {
      U8Carry carry = u8_carry_of(pw[j]);
      uint32_t em[4];
#pragma unroll
      for (int k = 0; k < 4; k++) {
        const uint32_t xn = (k < 3 ? w[j][(k + 1) & 3] : nw[j]) & 0x3F3F3F3Fu;
        const U8Word16 r = u8_to_utf16_word<true>(w[j][k], xn, carry);
        U[j][2 * k] = r.u01;
        U[j][2 * k + 1] = r.u23;
        em[k] = r.emit;
        flagged |= r.err;
      }
      M[j] = pack_emit(em[0], em[1], em[2], em[3]);
    }
    if (EDGE && !inside[j]) M[j] &= packed_inrange(in, g0 + (unsigned long long)j * 32u + lane);
    cnt[j] = (uint32_t)__popc(M[j]);
  }

  // ---- 3. warp scan, post the warp total, signal the scan warp ----
  uint32_t off[G];
  const uint32_t total = warp_exclusive_offsets<G>(cnt, off);
  if (lane == 0) sm.warp_total[par][warp] = total;
  __syncwarp();
  bar_arrive<kBarA>(par, kT16Threads);

  // ---- 4. compact into the warp's staging region ----
  uint16_t *stage = sm.stage[par][warp];
#pragma unroll
  for (int j = 0; j < G; j++) {
    uint16_t *sp = stage + 8 + off[j];
    const uint32_t m = M[j];
#pragma unroll
    for (int k = 0; k < 4; k++) {
#pragma unroll
      for (int b = 0; b < 4; b++) {
        const uint32_t reg = U[j][2 * k + (b >> 1)];
        const uint16_t unit = (uint16_t)((b & 1) ? (reg >> 16) : reg);
        if (m & (1u << (8 * b + 4 + k))) {
          *sp = unit;
          sp++;
        }
      }
    }
  }
  // exact error location (rare): the detector only says "somewhere in this granule or the 3 bytes before it"
  {
    bool bad = flagged != 0;
    if (EDGE) {
#pragma unroll
      for (int j = 0; j < G; j++) {
        const unsigned long long lo = (g0 + (unsigned long long)j * 32u + lane) * 16ull;
        if (lo < in.vend && in.vend <= lo + 16ull) bad = bad || tail_truncated16(in);
      }
    }
    if (bad) {
      // a thread's granules are 512 bytes apart: search each one's window separately
#pragma unroll
      for (int j = 0; j < G; j++) {
        const unsigned long long lo = (g0 + (unsigned long long)j * 32u + lane) * 16ull;
        u8_locate_error(in, scr, (long long)lo - 3, (long long)lo + 16);
      }
    }
  }
  __syncwarp();
  return total;
}

// Decoupled look-back, split in two so that a tile's aggregate is visible one whole tile period before anyone
// has to sum it:  publish_aggregate() as soon as the tile's count is known, lookback_wide() one tile later
// (32*KD predecessors per round; in steady state every one of them is already published, so nobody polls).
constexpr int kLookbackPerLane = 8;

__device__ __forceinline__ void publish_aggregate(unsigned long long *desc, uint32_t epoch, uint32_t tile,
                                                  unsigned long long agg) {
  if ((threadIdx.x & 31u) == 0)
    st_relaxed_u64(desc + tile, desc_pack(epoch, tile == 0 ? kStatusPrefix : kStatusAggregate, 0u, agg));
}

// Exclusive prefix of `tile` (sum of the aggregates of all tiles before it); upgrades the tile's descriptor to
// an inclusive prefix.  All 32 lanes of one warp.
__device__ __forceinline__ unsigned long long lookback_wide(unsigned long long *desc, uint32_t epoch, uint32_t tile,
                                                            unsigned long long agg) {
  constexpr int KD = kLookbackPerLane;
  const unsigned lane = threadIdx.x & 31u;
  if (tile == 0) return 0ull;
  unsigned long long sum = 0;
  long long base = (long long)tile - 1;
  while (true) {
    // lane l looks at tiles base-KD*l-k, k = 0..KD-1 (nearest first)
    unsigned long long d[KD];
#pragma unroll
    for (int k = 0; k < KD; k++) {
      const long long idx = base - (long long)(KD * lane) - k;
      d[k] = idx >= 0 ? ld_relaxed_u64(desc + idx) : desc_pack(epoch, kStatusPrefix, 0u, 0ull);
    }
#pragma unroll
    for (int k = 0; k < KD; k++) {
      const long long idx = base - (long long)(KD * lane) - k;
      while (desc_epoch(d[k]) != epoch || desc_status(d[k]) == 0) {  // start-up / stragglers only
        __nanosleep(256);
        d[k] = ld_relaxed_u64(desc + idx);
      }
    }
    unsigned long long mine = 0;
    bool have = false;
#pragma unroll
    for (int k = 0; k < KD; k++) {
      if (!have) mine += desc_value(d[k]);
      have = have || desc_status(d[k]) == kStatusPrefix;
    }
    const unsigned pm = __ballot_sync(kFull, have);
    const unsigned first = pm ? (unsigned)(__ffs((int)pm) - 1) : 32u;
    sum += warp_sum_u64(lane <= first ? mine : 0ull);
    if (pm) break;
    base -= 32 * KD;
  }
  if (lane == 0) st_relaxed_u64(desc + tile, desc_pack(epoch, kStatusPrefix, 0u, sum + agg));
  return sum;
}

template <int G, int MINB>
__global__ void __launch_bounds__(kT16Threads, MINB)
k_utf8_to_utf16(const char *ptr, size_t len, uint16_t *out, Scratch *scr, unsigned long long *desc, uint32_t epoch,
                uint32_t num_tiles, ResultPOD *res) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T16Smem<G> &sm = *reinterpret_cast<T16Smem<G> *>(smem_raw);
  const InView in = make_view16(ptr, len);
  constexpr unsigned long long kTileBytes = (unsigned long long)kCW * 32ull * G * 16ull;
  const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;

  // a CTA always holds the tickets of its current and its next tile
  if (threadIdx.x == 0) {
    const uint32_t t = atomicAdd(&scr->ticket, 2u);
    sm.first_tiles[0] = t;
    sm.first_tiles[1] = t + 1;
  }
  __syncthreads();
  uint32_t cur = sm.first_tiles[0], nxt = sm.first_tiles[1];

  if (warp < kCW) {
    uint32_t it = 0, prev_total = 0;
    while (cur < num_tiles) {
      const uint32_t par = it & 1u;
      const unsigned long long lo = (unsigned long long)cur * kTileBytes;
      uint32_t total;
      if (lo >= in.vbeg && lo + kTileBytes <= in.vend) total = transcode_tile<G, false>(in, scr, cur, par, sm);
      else total = transcode_tile<G, true>(in, scr, cur, par, sm);
      uint32_t next_tile;
      if (it > 0) {
        // ---- 5./6. previous tile: wait for its offsets, flush its staging region ----
        bar_sync<kBarB>(par ^ 1u, kT16Threads);
        const unsigned long long goff = sm.warp_goff[par ^ 1u][warp];
        next_tile = sm.tile_pub[par ^ 1u];  // published by the scan warp while it worked on the previous tile
        flush_warp(sm.stage[par ^ 1u][warp], out, goff, prev_total);
        __syncwarp();
      } else {
        next_tile = nxt;  // the CTA's second ticket
      }
      cur = next_tile;
      prev_total = total;
      it++;
    }
    if (it > 0) {
      const uint32_t par = (it - 1) & 1u;
      bar_sync<kBarB>(par, kT16Threads);
      flush_warp(sm.stage[par][warp], out, sm.warp_goff[par][warp], prev_total);
    }
  } else {
    // scan warp: iteration i publishes the aggregate of tile i, then resolves tile i-1 (published one tile ago)
    uint32_t it = 0, prev_tile = 0, prev_total = 0, prev_lane_excl = 0;
    while (cur < num_tiles) {
      const uint32_t par = it & 1u;
      uint32_t nt = 0;
      if (lane == 0) nt = atomicAdd(&scr->ticket, 1u);  // ticket of the tile after next
      bar_sync<kBarA>(par, kT16Threads);
      const uint32_t t = lane < kCW ? sm.warp_total[par][lane] : 0u;
      uint32_t incl = t;
#pragma unroll
      for (int o = 1; o < kCW; o <<= 1) {
        const uint32_t v = __shfl_up_sync(kFull, incl, o);
        if (lane >= (unsigned)o) incl += v;
      }
      const uint32_t tile_total = __shfl_sync(kFull, incl, kCW - 1);
      publish_aggregate(desc, epoch, cur, tile_total);
      if (it > 0) {
        const unsigned long long excl = lookback_wide(desc, epoch, prev_tile, prev_total);
        if (lane < kCW) sm.warp_goff[par ^ 1u][lane] = excl + prev_lane_excl;
        if (lane == 0) sm.tile_pub[par ^ 1u] = nxt;  // the tile the transcoding warps work on after `cur`
        __syncwarp();
        bar_arrive<kBarB>(par ^ 1u, kT16Threads);
      }
      prev_tile = cur;
      prev_total = tile_total;
      prev_lane_excl = incl - t;
      cur = nxt;
      nxt = __shfl_sync(kFull, nt, 0);
      it++;
    }
    if (it > 0) {
      const uint32_t par = (it - 1) & 1u;
      const unsigned long long excl = lookback_wide(desc, epoch, prev_tile, prev_total);
      if (lane < kCW) sm.warp_goff[par][lane] = excl + prev_lane_excl;
      if (lane == 0 && prev_tile == num_tiles - 1) st_relaxed_u64(&scr->acc0, excl + prev_total);
      __syncwarp();
      bar_arrive<kBarB>(par, kT16Threads);
    }
  }

  if (grid_last_thread(scr)) {
    const unsigned long long key = ld_relaxed_u64(&scr->err_key);
    if (key == kNoError) {
      res->error = kSuccess;
      res->reserved_ = 0;
      res->count = ld_relaxed_u64(&scr->acc0);
    } else {
      res->error = (int32_t)(key & 0xFFu);
      res->reserved_ = 0;
      res->count = key >> 8;
    }
    scratch_reset(scr);
  }
}

// Tuning knobs for experiments (tools/, profiles/): B200_TUNE_G = granules per lane (2..4),
// B200_TUNE_MINB = resident CTAs per SM the kernel is compiled for.
inline int env_int(const char *name, int lo, int hi, int dflt) {
  const char *e = getenv(name);
  if (!e || !*e) return dflt;
  const int v = atoi(e);
  return (v >= lo && v <= hi) ? v : dflt;
}
inline int tuned_g() {
  static int v = env_int("B200_TUNE_G", 2, 4, 4);
  return v;
}
inline int tuned_minb16() {
  static int v = env_int("B200_TUNE_MINB", 1, 4, 3);
  return v;
}

template <int G, int MINB>
cudaError_t launch_t16(const LaunchCtx &c, const char *in, size_t len, uint16_t *out, void *res, size_t tiles) {
  static int per_sm = 0;
  constexpr size_t smem = sizeof(T16Smem<G>);
  if (per_sm == 0) {
    cudaError_t e = cudaFuncSetAttribute(k_utf8_to_utf16<G, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int n = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k_utf8_to_utf16<G, MINB>, kT16Threads, smem);
    if (e != cudaSuccess) return e;
    per_sm = n < 1 ? 1 : n;
  }
  const size_t cap = (size_t)c.sm_count * per_sm;
  const unsigned grid = (unsigned)(tiles < cap ? tiles : cap);
  k_utf8_to_utf16<G, MINB><<<grid, kT16Threads, smem, c.stream>>>(in, len, out, c.scratch, c.desc, c.epoch,
                                                                 (uint32_t)tiles, static_cast<ResultPOD *>(res));
  count_launch(1);
  return cudaGetLastError();
}

inline size_t tiles16_for(const void *in, size_t len_bytes, int g) {
  const size_t span = (reinterpret_cast<uintptr_t>(in) & 15u) + len_bytes;
  const size_t gran = (span + 15) / 16;
  const size_t per_tile = (size_t)kCW * 32 * g;
  return (gran + per_tile - 1) / per_tile;
}

}  // namespace

size_t utf8_to_utf16_tiles(const void *in, size_t len) { return tiles16_for(in, len, tuned_g()); }

cudaError_t launch_convert_utf8_to_utf16le(const LaunchCtx &c, const char *in, size_t len, uint16_t *out, void *res) {
  const int g = tuned_g();
  const size_t tiles = tiles16_for(in, len, g);
  if (tiles > c.desc_capacity || tiles > 0xFFFFFFF0ull) return cudaErrorInvalidValue;
  const int mb = tuned_minb16();
  switch (g) {
    case 2:
      if (mb == 2) return launch_t16<2, 2>(c, in, len, out, res, tiles);
      if (mb == 4) return launch_t16<2, 4>(c, in, len, out, res, tiles);
      return launch_t16<2, 3>(c, in, len, out, res, tiles);
    case 3:
      if (mb == 2) return launch_t16<3, 2>(c, in, len, out, res, tiles);
      return launch_t16<3, 3>(c, in, len, out, res, tiles);
    default:
      if (mb == 2) return launch_t16<4, 2>(c, in, len, out, res, tiles);
      return launch_t16<4, 3>(c, in, len, out, res, tiles);
  }
}

}  // namespace b200
