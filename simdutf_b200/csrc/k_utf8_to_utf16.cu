// k_utf8_to_utf16.cu — sm_100a kernels K3a/K3b: convert_utf8_to_utf16le[_with_errors]
// (reference include/simdutf/implementation.h:3709-3745; semantics src/scalar/utf8_to_utf16/utf8_to_utf16.h:128-255).
//
// Two phases, both pure streaming kernels made of independent warps (no CTA barrier in the hot loop, no
// inter-CTA wait):
//
//   K3a  k_utf16_tile_counts   per warp-tile (32*G granules = 512*G contiguous input bytes) the number of UTF-16
//        units it will produce — the popcount reduction of utf16_length_from_utf8 (reference
//        src/scalar/utf8.h:243-255) kept per tile — plus one total per chunk of 64 warp-tiles; the CTA that
//        finishes last turns the chunk totals into exclusive chunk offsets (and the grand total).
//        HBM-bound: reads the input once, writes 2 bytes per KiB.
//   K3b  k_utf8_to_utf16_emit  every warp: output offset of its tile = chunk offset + the counts of the tiles
//        before it in the chunk (one masked warp reduction); coalesced 128-bit loads; branch-free SWAR
//        transcoder + validation detector (swar.h: u8_to_utf16_word) with the results kept in registers;
//        popcount + warp scan; compaction into the warp's own shared-memory staging region, laid out so that
//        its 16-byte vectors line up with 16-byte-aligned output addresses; 128-bit streaming stores.
//        The next tile's granules are fetched while the current one is compacted and stored.
//        ALU-pipe-bound (the per-byte SWAR work); the input comes from HBM a second time.  DESIGN.md explains
//        why a one-pass chained scan lost to this on B200: with ~600 resident tiles every look-back waits
//        for the slowest of its predecessors, and that convoy costs more than one extra streaming read.
//
// Emit rule and values: swar.h (one unit per non-continuation byte + one for the byte after a byte >= 0xF0, so
// an output buffer of utf16_length_from_utf8() units is never overrun, even for invalid input).
#include <cstdlib>

#include "device_common.cuh"
#include "launch.h"

namespace b200 {

namespace {

constexpr int kWarpsPerCta = 8;
constexpr int kThreads = kWarpsPerCta * 32;
constexpr uint32_t kChunkTiles = 64;  // warp-tiles per chunk (one chunk total / chunk offset)

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

// A tile is "interior" when all its granules lie inside the buffer (the words on either side are always
// fetched with guarded loads).
template <int G>
__device__ __forceinline__ bool tile_is_interior(const InView &in, unsigned long long g0) {
  const unsigned long long lo = g0 * 16ull, hi = (g0 + 32ull * G) * 16ull;
  return lo >= in.vbeg && hi <= in.vend;
}

// Loads the G granules of this lane (granule g0 + j*32 + lane).
template <int G, bool EDGE>
__device__ __forceinline__ void load_tile(const InView &in, unsigned long long g0, uint32_t (&w)[G][4], bool (&inside)[G]) {
  const unsigned lane = threadIdx.x & 31u;
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
}

// ---------------------------------------------------------------------------------------------
// K3a: per-tile unit counts
// ---------------------------------------------------------------------------------------------
template <int G, bool EDGE>
__device__ __forceinline__ uint32_t count_tile(const InView &in, unsigned long long g0) {
  const unsigned lane = threadIdx.x & 31u;
  uint32_t w[G][4];
  bool inside[G];
  load_tile<G, EDGE>(in, g0, w, inside);
  uint32_t cnt = 0;
#pragma unroll
  for (int j = 0; j < G; j++) {
    // the word before the granule: only its last byte matters (>= 0xF0 makes byte 0 emit a low surrogate)
    const uint32_t give = (lane == 31 && j > 0) ? w[j - 1][3] : w[j][3];
    uint32_t pw = __shfl_sync(kFull, give, (lane + 31u) & 31u);
    if (j == 0 && lane == 0) pw = load_word_guarded(in, (long long)(g0 * 4ull) - 1);
    uint32_t em[4];
    u8_emit16_masks(w[j], pw, em);
    uint32_t m = pack_emit(em[0], em[1], em[2], em[3]);
    if (EDGE && !inside[j]) m &= packed_inrange(in, g0 + (unsigned long long)j * 32u + lane);
    cnt += (uint32_t)__popc(m);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(kFull, cnt, o);
  return cnt;
}

template <int G>
__global__ void __launch_bounds__(kThreads) k_utf16_tile_counts(const char *ptr, size_t len, uint16_t *tile_cnt,
                                                                 unsigned long long *chunk_off, uint32_t num_tiles,
                                                                 uint32_t num_chunks, Scratch *scr) {
  __shared__ uint32_t s_tot[kWarpsPerCta];
  __shared__ unsigned long long s_warp_sum[kWarpsPerCta];
  __shared__ unsigned long long s_carry;
  __shared__ bool s_last;
  const InView in = make_view16(ptr, len);
  const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  for (uint32_t chunk = blockIdx.x; chunk < num_chunks; chunk += gridDim.x) {
    uint32_t mine = 0;
    for (uint32_t i = warp; i < kChunkTiles; i += kWarpsPerCta) {
      const uint32_t t = chunk * kChunkTiles + i;
      if (t >= num_tiles) break;
      const unsigned long long g0 = (unsigned long long)t * (32ull * G);
      const uint32_t c = tile_is_interior<G>(in, g0) ? count_tile<G, false>(in, g0) : count_tile<G, true>(in, g0);
      if (lane == 0) tile_cnt[t] = (uint16_t)c;
      mine += c;
    }
    if (lane == 0) s_tot[warp] = mine;
    __syncthreads();
    if (threadIdx.x == 0) {
      uint32_t tot = 0;
#pragma unroll
      for (int k = 0; k < kWarpsPerCta; k++) tot += s_tot[k];
      chunk_off[chunk] = tot;  // turned into an exclusive offset below
    }
    __syncthreads();
  }
  // the CTA that finishes last scans the chunk totals (num_chunks is small: 16 Ki per GiB of input at G = 2)
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    s_last = atomicAdd(&scr->done, 1u) == gridDim.x - 1;
    s_carry = 0;
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  for (uint32_t base = 0; base < num_chunks; base += kThreads) {
    const uint32_t i = base + threadIdx.x;
    const unsigned long long v = i < num_chunks ? ld_relaxed_u64(chunk_off + i) : 0ull;
    unsigned long long incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned long long t = __shfl_up_sync(kFull, incl, o);
      if (lane >= (unsigned)o) incl += t;
    }
    if (lane == 31) s_warp_sum[warp] = incl;
    __syncthreads();
    unsigned long long before = s_carry;
#pragma unroll
    for (int k = 0; k < kWarpsPerCta; k++)
      if ((unsigned)k < warp) before += s_warp_sum[k];
    if (i < num_chunks) chunk_off[i] = before + incl - v;
    __syncthreads();
    if (threadIdx.x == kThreads - 1) s_carry = before + incl;
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    chunk_off[num_chunks] = s_carry;  // grand total
    scr->done = 0;
    __threadfence();
  }
}

// ---------------------------------------------------------------------------------------------
// K3b: transcode + emit
// ---------------------------------------------------------------------------------------------
template <int G>
struct EmitSmem {
  static constexpr uint32_t kWarpUnits = 32u * G * 16u;  // a warp emits at most one unit per input byte
  static constexpr uint32_t kRegion = kWarpUnits + 16u;  // up to 7 units of alignment padding in front + vector tail
  alignas(16) uint16_t stage[kWarpsPerCta][kRegion];
};

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

// The part of a tile that needs the input words: leaves candidate units and emit masks in registers.
template <int G, bool EDGE>
__device__ __forceinline__ void transcode_words(const InView &in, unsigned long long g0, const uint32_t (&w)[G][4],
                                                const bool (&inside)[G], uint32_t (&U)[G][8], uint32_t (&M)[G],
                                                uint32_t &flagged) {
  const unsigned lane = threadIdx.x & 31u;
  uint32_t pw[G], nw[G];
  neighbour_words<G>(in, g0, w, pw, nw);
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
  }
}

template <int G, int MINB>
__global__ void __launch_bounds__(kThreads, MINB)
k_utf8_to_utf16_emit(const char *ptr, size_t len, uint16_t *out, const uint16_t *tile_cnt,
                     const unsigned long long *chunk_off, uint32_t num_tiles, uint32_t num_chunks, Scratch *scr,
                     ResultPOD *res) {
  __shared__ EmitSmem<G> sm;
  const InView in = make_view16(ptr, len);
  const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  const uint32_t nwarps = gridDim.x * kWarpsPerCta;
  uint16_t *stage = sm.stage[warp];

  uint32_t tile = blockIdx.x * kWarpsPerCta + warp;
  uint32_t w[G][4];
  bool inside[G];
  bool interior = false;
  if (tile < num_tiles) {
    const unsigned long long g0 = (unsigned long long)tile * (32ull * G);
    interior = tile_is_interior<G>(in, g0);
    if (interior) load_tile<G, false>(in, g0, w, inside);
    else load_tile<G, true>(in, g0, w, inside);
  }
  while (tile < num_tiles) {
    const unsigned long long g0 = (unsigned long long)tile * (32ull * G);
    // ---- where the tile's units go: chunk offset + counts of the chunk's earlier tiles ----
    const uint32_t chunk = tile / kChunkTiles, in_chunk = tile % kChunkTiles;
    uint32_t before = 0;
    {
      const uint16_t *c = tile_cnt + (size_t)chunk * kChunkTiles;
      if (lane < in_chunk) before += c[lane];
      if (lane + 32u < in_chunk) before += c[lane + 32u];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) before += __shfl_xor_sync(kFull, before, o);
    }
    const unsigned long long goff = chunk_off[chunk] + before;

    // ---- transcode into registers ----
    uint32_t U[G][8], M[G], cnt[G], off[G];
    uint32_t flagged = 0;
    if (interior) transcode_words<G, false>(in, g0, w, inside, U, M, flagged);
    else transcode_words<G, true>(in, g0, w, inside, U, M, flagged);
    const bool was_interior = interior;

    // ---- fetch the next tile while this one is compacted and stored ----
    const uint32_t next = tile + nwarps;
    if (next < num_tiles) {
      const unsigned long long n0 = (unsigned long long)next * (32ull * G);
      interior = tile_is_interior<G>(in, n0);
      if (interior) load_tile<G, false>(in, n0, w, inside);
      else load_tile<G, true>(in, n0, w, inside);
    }

#pragma unroll
    for (int j = 0; j < G; j++) cnt[j] = (uint32_t)__popc(M[j]);
    const uint32_t total = warp_exclusive_offsets<G>(cnt, off);

    // ---- compact into the staging region; unit i sits at stage[a + i], a = misalignment of the destination ----
    uint16_t *dst = out + goff;
    const uint32_t a = (uint32_t)(reinterpret_cast<uintptr_t>(dst) >> 1) & 7u;
#pragma unroll
    for (int j = 0; j < G; j++) {
      uint16_t *sp = stage + a + off[j];
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
      if (!was_interior) {
#pragma unroll
        for (int j = 0; j < G; j++) {
          const unsigned long long lo = (g0 + (unsigned long long)j * 32u + lane) * 16ull;
          if (lo < in.vend && in.vend <= lo + 16ull) bad = bad || tail_truncated16(in);
        }
      }
      if (bad) {
#pragma unroll
        for (int j = 0; j < G; j++) {
          const unsigned long long lo = (g0 + (unsigned long long)j * 32u + lane) * 16ull;
          u8_locate_error(in, scr, (long long)lo - 3, (long long)lo + 16);
        }
      }
    }
    __syncwarp();

    // ---- staging -> global: vector v of the region is vector v of the 16-byte-aligned destination ----
    {
      const uint32_t nvec = (a + total + 7u) >> 3;
      uint16_t *dbase = dst - a;
      const uint4 *sv = reinterpret_cast<const uint4 *>(stage);
      for (uint32_t v = lane; v < nvec; v += 32u) {
        const bool full = (v > 0 || a == 0) && (8u * v + 8u <= a + total);
        if (full) {
          stg_stream_v4(reinterpret_cast<uint4 *>(dbase + 8u * v), sv[v]);
        } else {
#pragma unroll
          for (uint32_t t = 0; t < 8; t++) {
            const uint32_t e = 8u * v + t;
            if (e >= a && e < a + total) dbase[e] = stage[e];
          }
        }
      }
    }
    __syncwarp();  // the region is rewritten by the next tile
    tile = next;
  }

  if (grid_last_thread(scr)) {
    const unsigned long long key = ld_relaxed_u64(&scr->err_key);
    if (key == kNoError) {
      res->error = kSuccess;
      res->reserved_ = 0;
      res->count = chunk_off[num_chunks];
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
  static int v = env_int("B200_TUNE_MINB", 2, 4, 3);
  return v;
}

inline size_t tiles16_for(const void *in, size_t len_bytes, int g) {
  const size_t span = (reinterpret_cast<uintptr_t>(in) & 15u) + len_bytes;
  const size_t gran = (span + 15) / 16;
  const size_t per_tile = (size_t)32 * g;
  return (gran + per_tile - 1) / per_tile;
}

template <int G, int MINB>
cudaError_t launch_t16(const LaunchCtx &c, const char *in, size_t len, uint16_t *out, void *res, size_t tiles) {
  static int per_sm_emit = 0;
  if (per_sm_emit == 0) {
    int n = 0;
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k_utf8_to_utf16_emit<G, MINB>, kThreads, 0);
    if (e != cudaSuccess) return e;
    per_sm_emit = n < 1 ? 1 : n;
  }
  const size_t chunks = (tiles + kChunkTiles - 1) / kChunkTiles;
  // workspace carved out of the descriptor array: [chunks + 1] u64 chunk offsets, then one u16 per tile
  unsigned long long *chunk_off = c.desc;
  uint16_t *tile_cnt = reinterpret_cast<uint16_t *>(c.desc + chunks + 1);
  {
    const size_t cap = (size_t)c.sm_count * 8;
    const unsigned grid = (unsigned)(chunks < cap ? chunks : cap);
    k_utf16_tile_counts<G><<<grid, kThreads, 0, c.stream>>>(in, len, tile_cnt, chunk_off, (uint32_t)tiles, (uint32_t)chunks,
                                                           c.scratch);
  }
  {
    const size_t ctas = (tiles + kWarpsPerCta - 1) / kWarpsPerCta;
    const size_t cap = (size_t)c.sm_count * per_sm_emit;
    const unsigned grid = (unsigned)(ctas < cap ? ctas : cap);
    k_utf8_to_utf16_emit<G, MINB><<<grid, kThreads, 0, c.stream>>>(in, len, out, tile_cnt, chunk_off, (uint32_t)tiles,
                                                                  (uint32_t)chunks, c.scratch, static_cast<ResultPOD *>(res));
  }
  count_launch(2);
  return cudaGetLastError();
}

}  // namespace

// Workspace, in 8-byte descriptor slots, the two kernels need for an input of `len` bytes.
size_t utf8_to_utf16_tiles(const void *in, size_t len) {
  const size_t tiles = tiles16_for(in, len, tuned_g());
  const size_t chunks = (tiles + kChunkTiles - 1) / kChunkTiles;
  return chunks + 1 + (tiles * sizeof(uint16_t) + 7) / 8 + 1;
}

cudaError_t launch_convert_utf8_to_utf16le(const LaunchCtx &c, const char *in, size_t len, uint16_t *out, void *res) {
  const int g = tuned_g();
  const size_t tiles = tiles16_for(in, len, g);
  if (utf8_to_utf16_tiles(in, len) > c.desc_capacity || tiles > 0xFFFFFF00ull) return cudaErrorInvalidValue;
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
