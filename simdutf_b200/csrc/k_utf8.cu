// k_utf8.cu — sm_100a kernels whose input is UTF-8:
//   K1  validate_utf8_with_errors          (reference src/scalar/utf8.h:102-200)
//   K2  count_utf8 / utf16_length_from_utf8 (reference src/scalar/utf8.h:230-255)
//   K3  convert_utf8_to_utf16le[_with_errors] (reference src/scalar/utf8_to_utf16/utf8_to_utf16.h:128-255)
//   K4  convert_utf8_to_utf32[_with_errors]   (reference src/scalar/utf8_to_utf32/utf8_to_utf32.h:106-212)
//
// Data layout: the input is read once, as 16-byte granules with fully coalesced 128-bit streaming loads
// (lane l of a warp owns granule g0 + j*32 + l, j = 0..ITEMS-1, so one load instruction covers 512
// contiguous bytes).  The 3-byte look-behind / look-ahead a byte's verdict and value depend on come from
// the neighbouring lane by warp shuffle; only the two words flanking a warp's 2 KiB chunk are re-read from
// L2.  Validation and counting are reductions (first error = atomicMin of position<<8|code).  Transcoding
// is one pass: per-granule output counts -> block scan -> decoupled look-back across tiles (device_common.cuh)
// -> units staged in shared memory at their final offsets -> 16-byte coalesced streaming stores.
#include <cstdlib>

#include "bitplane.h"
#include "device_common.cuh"
#include "launch.h"

namespace b200 {

namespace {

__device__ __forceinline__ InView make_view(const void *p, size_t len_bytes) {
  InView v;
  const uintptr_t a = reinterpret_cast<uintptr_t>(p);
  v.base = reinterpret_cast<const uint4 *>(a & ~uintptr_t(15));
  v.vbeg = a & 15u;
  v.vend = v.vbeg + len_bytes;
  return v;
}

// True iff the last three bytes of the buffer start a sequence that the end of the buffer cuts short.
__device__ __forceinline__ bool tail_truncated(const InView &in) {
  const long long e = (long long)in.vend;
  const uint8_t *p = reinterpret_cast<const uint8_t *>(in.base);
  const uint32_t b1 = p[e - 1];
  const uint32_t b2 = (e - 2 >= (long long)in.vbeg) ? p[e - 2] : 0u;
  const uint32_t b3 = (e - 3 >= (long long)in.vbeg) ? p[e - 3] : 0u;
  return u8_incomplete_tail(b1, b2, b3);
}

// Validation of the ITEMS granules a thread holds.  Flags a granule with the SWAR detector, then pins the
// exact (code, position) with u8_locate_error on [lo-3, hi).  The granule that contains the last byte of
// the buffer also checks for a sequence cut short by the end of the buffer.
template <int ITEMS>
__device__ __forceinline__ void validate_items(const InView &in, Scratch *scr, unsigned long long g0,
                                               const uint32_t (&w)[ITEMS][4], const uint32_t (&pw)[ITEMS]) {
  const unsigned lane = threadIdx.x & 31u;
#pragma unroll
  for (int j = 0; j < ITEMS; j++) {
    const unsigned long long g = g0 + (unsigned long long)j * 32u + lane;
    const unsigned long long lo = g * 16ull;
    const uint32_t any_hi = (w[j][0] | w[j][1] | w[j][2] | w[j][3] | pw[j]) & kH;
    bool flagged = false;
    if (any_hi) flagged = u8_check_granule(w[j], pw[j]) != 0;
    if (lo < in.vend && in.vend <= lo + 16ull) flagged = flagged || tail_truncated(in);  // holds the final byte
    if (flagged) u8_locate_error(in, scr, (long long)lo - 3, (long long)lo + 16);
  }
}

__device__ __forceinline__ void write_result_from_key(ResultPOD *res, unsigned long long key,
                                                      unsigned long long success_count) {
  if (key == kNoError) {
    res->error = kSuccess;
    res->reserved_ = 0;
    res->count = success_count;
  } else {
    res->error = (int32_t)(key & 0xFFu);
    res->reserved_ = 0;
    res->count = key >> 8;
  }
}

// ---------------------------------------------------------------------------------------------
// K1: validate_utf8_with_errors
// ---------------------------------------------------------------------------------------------
template <int ITEMS>
__global__ void __launch_bounds__(kBlock) k_validate_utf8(const char *ptr, size_t len, Scratch *scr, ResultPOD *res) {
  static_assert(ITEMS % 2 == 0, "a bit-plane block is two granules");
  __shared__ uint4 s_slab[kWarps][32 * ITEMS + 4 * ITEMS];
  const InView in = make_view(ptr, len);
  const unsigned lane = threadIdx.x & 31u;
  const unsigned long long ngran = (in.vend + 15ull) >> 4;
  const unsigned long long chunk_gran = 32ull * ITEMS;
  const unsigned long long nchunks = (ngran + chunk_gran - 1) / chunk_gran;
  const unsigned long long nwarps = (unsigned long long)gridDim.x * kWarps;
  for (unsigned long long chunk = (unsigned long long)blockIdx.x * kWarps + (threadIdx.x >> 5); chunk < nchunks;
       chunk += nwarps) {
    const unsigned long long g0 = chunk * chunk_gran;
    uint32_t w[ITEMS][4];
    bool inside[ITEMS];
#pragma unroll
    for (int j = 0; j < ITEMS; j++) load_granule(in, g0 + (unsigned long long)j * 32u + lane, w[j], inside[j]);
    uint32_t hi = 0;
#pragma unroll
    for (int j = 0; j < ITEMS; j++) hi |= w[j][0] | w[j][1] | w[j][2] | w[j][3];
    // Word that precedes the chunk: a lead there may still be waiting for continuation bytes.
    uint32_t before = 0;
    if (lane == 0) before = load_word_guarded(in, (long long)(g0 * 4ull) - 1);
    const bool last_chunk = chunk == nchunks - 1;
    // All-ASCII fast path (config 1): nothing to verify unless the buffer ends here.
    if (!last_chunk && !__any_sync(kFull, ((hi | before) & kH) != 0u)) continue;
    // Non-ASCII chunk: re-lay the chunk out so that every lane holds 16*ITEMS CONTIGUOUS bytes (through the warp's
    // shared-memory slab; granule g sits at 16 * (g + g/8) so that both the coalesced-order stores and the
    // lane-contiguous loads are bank-conflict free), then validate in bit-plane form (bitplane.h): one transposition
    // and ~30 bitwise instructions per 32 bytes.  A flagged block is pinned down exactly by u8_locate_error.
    uint4 *slab = s_slab[threadIdx.x >> 5];
    __syncwarp();
#pragma unroll
    for (int j = 0; j < ITEMS; j++) {
      const uint32_t g = (uint32_t)j * 32u + lane;
      slab[g + (g >> 3)] = make_uint4(w[j][0], w[j][1], w[j][2], w[j][3]);
    }
    __syncwarp();
    uint32_t B[ITEMS / 2][8];
#pragma unroll
    for (int j = 0; j < ITEMS; j++) {
      const uint32_t g = lane * (uint32_t)ITEMS + (uint32_t)j;
      const uint4 v = slab[g + (g >> 3)];
      B[j >> 1][4 * (j & 1) + 0] = v.x;
      B[j >> 1][4 * (j & 1) + 1] = v.y;
      B[j >> 1][4 * (j & 1) + 2] = v.z;
      B[j >> 1][4 * (j & 1) + 3] = v.w;
    }
    uint32_t pword = __shfl_up_sync(kFull, B[ITEMS / 2 - 1][7], 1);
    if (lane == 0) pword = before;
    bp::VCarry vc = bp::vcarry_from_word(pword);
    uint32_t badblocks = 0;
#pragma unroll
    for (int j = 0; j < ITEMS / 2; j++) {
      bp::transpose_in(B[j]);
      if (bp::utf8_check_block(B[j], vc)) badblocks |= 1u << j;
    }
    const unsigned long long r0 = (g0 + (unsigned long long)lane * ITEMS) * 16ull;  // first byte of this lane's region
    if (last_chunk) {
#pragma unroll
      for (int j = 0; j < ITEMS / 2; j++) {
        const unsigned long long b0 = r0 + 32ull * j;
        if (b0 < in.vend && in.vend <= b0 + 32ull && tail_truncated(in)) badblocks |= 1u << j;
      }
    }
    if (badblocks) {
#pragma unroll
      for (int j = 0; j < ITEMS / 2; j++) {
        const long long b0 = (long long)(r0 + 32ull * j);
        if (badblocks & (1u << j)) u8_locate_error(in, scr, b0 - 3, b0 + 32);
      }
    }
  }
  if (grid_last_thread(scr)) {
    write_result_from_key(res, ld_relaxed_u64(&scr->err_key), len);
    scratch_reset(scr);
  }
}

// ---------------------------------------------------------------------------------------------
// K2: count_utf8 (MODE 0) / utf16_length_from_utf8 (MODE 1).  Never validates.
// ---------------------------------------------------------------------------------------------
template <int MODE, int ITEMS>
__global__ void __launch_bounds__(kBlock) k_count_utf8(const char *ptr, size_t len, Scratch *scr,
                                                       unsigned long long *out) {
  __shared__ unsigned long long s_part[kWarps];
  const InView in = make_view(ptr, len);
  const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  const unsigned long long ngran = (in.vend + 15ull) >> 4;
  const unsigned long long chunk_gran = 32ull * ITEMS;
  const unsigned long long nchunks = (ngran + chunk_gran - 1) / chunk_gran;
  const unsigned long long nwarps = (unsigned long long)gridDim.x * kWarps;
  unsigned long long total = 0;
  for (unsigned long long chunk = (unsigned long long)blockIdx.x * kWarps + warp; chunk < nchunks; chunk += nwarps) {
    const unsigned long long g0 = chunk * chunk_gran;
    uint32_t w[ITEMS][4];
    bool inside[ITEMS];
#pragma unroll
    for (int j = 0; j < ITEMS; j++) load_granule(in, g0 + (unsigned long long)j * 32u + lane, w[j], inside[j]);
    uint32_t cnt = 0;
#pragma unroll
    for (int j = 0; j < ITEMS; j++) {
      const unsigned long long g = g0 + (unsigned long long)j * 32u + lane;
#pragma unroll
      for (int k = 0; k < 4; k++) {
        uint32_t m = u8_noncont(w[j][k]);
        if (MODE == 1) m |= u8_ge_f0(w[j][k]) >> 1;  // bit 6 of the same byte: one popc counts both
        if (!inside[j]) {
          const uint32_t r = inrange_mask_word(in, g, k);
          m &= r | (r >> 1);
        }
        cnt += (uint32_t)__popc(m);
      }
    }
    total += cnt;
  }
  total = warp_sum_u64(total);
  if (lane == 0) s_part[warp] = total;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long t = 0;
#pragma unroll
    for (int i = 0; i < kWarps; i++) t += s_part[i];
    if (t) atomicAdd(&scr->acc0, t);
  }
  if (grid_last_thread(scr)) {
    *out = ld_relaxed_u64(&scr->acc0);
    scratch_reset(scr);
  }
}

// ---------------------------------------------------------------------------------------------
// K3 / K4: UTF-8 -> UTF-16LE (OutT = uint16_t) / UTF-32 (OutT = uint32_t), validating.
// Persistent CTAs take tiles of kBlock*ITEMS granules from an atomic ticket (in order: required by the
// look-back scan).
// ---------------------------------------------------------------------------------------------
template <typename OutT, int ITEMS>
struct ConvertSmem {
  static constexpr uint32_t kTileBytes = kBlock * ITEMS * 16;
  static constexpr uint32_t kPad = 16 / sizeof(OutT);
  alignas(16) OutT out[kTileBytes + kPad];  // at most one element per input byte
  uint32_t warp_tot[kWarps];
  uint32_t tile;
  unsigned long long excl;
};

// One tile.  EDGE = the tile touches the first or last byte of the buffer (granules may be partial: guarded
// loads, in-range masks, truncated-tail check); interior tiles compile all of that away.
template <typename OutT, int ITEMS, bool VALIDATE, bool EDGE>
__device__ __forceinline__ void convert_tile(const InView &in, OutT *out, Scratch *scr, unsigned long long *desc,
                                             uint32_t epoch, uint32_t num_tiles, uint32_t tile,
                                             ConvertSmem<OutT, ITEMS> &sm) {
  constexpr bool k16 = sizeof(OutT) == 2;
  const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  const unsigned long long g0 = ((unsigned long long)tile * kWarps + warp) * (32ull * ITEMS);

  // ---- load + neighbours -------------------------------------------------------------------
  uint32_t w[ITEMS][4];
  bool inside[ITEMS];
#pragma unroll
  for (int j = 0; j < ITEMS; j++) {
    const unsigned long long g = g0 + (unsigned long long)j * 32u + lane;
    if (EDGE) {
      load_granule(in, g, w[j], inside[j]);
    } else {
      const uint4 v = ldg_stream_v4(in.base + g);
      w[j][0] = v.x; w[j][1] = v.y; w[j][2] = v.z; w[j][3] = v.w;
      inside[j] = true;
    }
  }
  uint32_t pw[ITEMS], nw[ITEMS];
  neighbour_words<ITEMS>(in, g0, w, pw, nw);

  // ---- per-granule output counts (phase 1: cheap SWAR popcounts, nothing is kept but the counts) ------
  uint32_t cnt[ITEMS], off[ITEMS];
#pragma unroll
  for (int j = 0; j < ITEMS; j++) {
    uint32_t em[4];
    if (k16) u8_emit16_masks(w[j], pw[j], em);
    else u8_emit32_masks(w[j], em);
    if (EDGE && !inside[j]) {
      const unsigned long long g = g0 + (unsigned long long)j * 32u + lane;
#pragma unroll
      for (int k = 0; k < 4; k++) em[k] &= inrange_mask_word(in, g, k);
    }
    cnt[j] = (uint32_t)(__popc(em[0]) + __popc(em[1]) + __popc(em[2]) + __popc(em[3]));
  }
  const uint32_t tile_total = block_exclusive_offsets<ITEMS>(cnt, off, sm.warp_tot);

  // ---- publish aggregate, look back for this tile's output offset ---------------------------
  if (warp == 0) {
    unsigned long long excl;
    uint32_t aux;
    tile_lookback(desc, epoch, tile, tile_total, 0u, excl, aux);
    if (lane == 0) {
      sm.excl = excl;
      if (tile == num_tiles - 1) st_relaxed_u64(&scr->acc0, excl + tile_total);
    }
  }
  __syncthreads();
  const unsigned long long excl = sm.excl;
  OutT *gdst = out + excl;
  const uint32_t shift = staging_shift(gdst);

  // ---- phase 2: decode (+ validate) into the staging buffer at final (tile-relative) offsets ----------
  if (k16) {
#pragma unroll
    for (int j = 0; j < ITEMS; j++) {
      const unsigned long long g = g0 + (unsigned long long)j * 32u + lane;
      OutT *sp = sm.out + shift + off[j];
      U8Carry carry = u8_carry_of(pw[j]);
      uint32_t flagged = 0;
#pragma unroll
      for (int k = 0; k < 4; k++) {
        const uint32_t xn = (k < 3 ? w[j][(k + 1) & 3] : nw[j]) & 0x3F3F3F3Fu;
        const U8Word16 r = u8_to_utf16_word<VALIDATE>(w[j][k], xn, carry);
        uint32_t em = r.emit;
        if (EDGE && !inside[j]) em &= inrange_mask_word(in, g, k);
        flagged |= r.err;
        if (em & 0x00000080u) *sp++ = (OutT)(r.u01 & 0xFFFFu);
        if (em & 0x00008000u) *sp++ = (OutT)(r.u01 >> 16);
        if (em & 0x00800000u) *sp++ = (OutT)(r.u23 & 0xFFFFu);
        if (em & 0x80000000u) *sp++ = (OutT)(r.u23 >> 16);
      }
      if (VALIDATE) {
        const unsigned long long lo = g * 16ull;
        bool bad = flagged != 0;
        if (EDGE && lo < in.vend && in.vend <= lo + 16ull) bad = bad || tail_truncated(in);
        if (bad) u8_locate_error(in, scr, (long long)lo - 3, (long long)lo + 16);
      }
    }
  } else {
    if (VALIDATE) validate_items<ITEMS>(in, scr, g0, w, pw);
#pragma unroll
    for (int j = 0; j < ITEMS; j++) {
      uint32_t em[4];
      u8_emit32_masks(w[j], em);
      if (EDGE && !inside[j]) {
        const unsigned long long g = g0 + (unsigned long long)j * 32u + lane;
#pragma unroll
        for (int k = 0; k < 4; k++) em[k] &= inrange_mask_word(in, g, k);
      }
      uint32_t o = shift + off[j];
      u8_emit32_granule(w[j], nw[j], em, [&](uint32_t u) { sm.out[o++] = (OutT)u; });
    }
  }
  __syncthreads();
  copy_out_aligned<OutT>(sm.out, gdst, shift, tile_total);
  __syncthreads();  // staging buffer and sm.tile are reused by the next tile
}

template <typename OutT, int ITEMS, bool VALIDATE, int MINB>
__global__ void __launch_bounds__(kBlock, MINB) k_convert_utf8(const char *ptr, size_t len, OutT *out, Scratch *scr,
                                                               unsigned long long *desc, uint32_t epoch,
                                                               uint32_t num_tiles, ResultPOD *res) {
  __shared__ ConvertSmem<OutT, ITEMS> sm;
  const InView in = make_view(ptr, len);
  constexpr unsigned long long kTileBytes = (unsigned long long)kBlock * ITEMS * 16ull;

  while (true) {
    if (threadIdx.x == 0) sm.tile = atomicAdd(&scr->ticket, 1u);
    __syncthreads();
    const uint32_t tile = sm.tile;
    if (tile >= num_tiles) break;
    const unsigned long long lo = (unsigned long long)tile * kTileBytes;
    if (lo >= in.vbeg && lo + kTileBytes <= in.vend) {
      convert_tile<OutT, ITEMS, VALIDATE, false>(in, out, scr, desc, epoch, num_tiles, tile, sm);
    } else {
      convert_tile<OutT, ITEMS, VALIDATE, true>(in, out, scr, desc, epoch, num_tiles, tile, sm);
    }
  }

  if (grid_last_thread(scr)) {
    write_result_from_key(res, ld_relaxed_u64(&scr->err_key), ld_relaxed_u64(&scr->acc0));
    scratch_reset(scr);
  }
}

__global__ void k_write_result(ResultPOD *res, int32_t error, unsigned long long count) {
  res->error = error;
  res->reserved_ = 0;
  res->count = count;
}
__global__ void k_write_full_result(FullResultPOD *res, int32_t error, unsigned long long in_count,
                                    unsigned long long out_count) {
  res->error = error;
  res->reserved_ = 0;
  res->input_count = in_count;
  res->output_count = out_count;
}
__global__ void k_write_u64(unsigned long long *dst, unsigned long long v) { *dst = v; }
__global__ void k_scratch_init(Scratch *scr) { scratch_reset(scr); }

constexpr int kStreamItems = 4;   // granules per thread per iteration in the reduction kernels
constexpr int kConv16Items = 4;   // 16 KiB tiles, 32 KiB staging
constexpr int kConv32Items = 2;   //  8 KiB tiles, 32 KiB staging

inline unsigned reduction_grid(const LaunchCtx &c, size_t len_bytes, int items) {
  const unsigned long long chunks = (len_bytes + 16 + 511ull * items) / (512ull * items);
  const unsigned long long ctas = (chunks + kWarps - 1) / kWarps;
  const unsigned long long cap = (unsigned long long)c.sm_count * 8;  // 8 CTAs x 256 threads = full SM
  return (unsigned)(ctas < 1 ? 1 : (ctas < cap ? ctas : cap));
}

inline size_t tiles_for(const void *in, size_t len_bytes, int items) {
  const size_t span = (reinterpret_cast<uintptr_t>(in) & 15u) + len_bytes;
  const size_t gran = (span + 15) / 16;
  const size_t per_tile = (size_t)kBlock * items;
  return (gran + per_tile - 1) / per_tile;
}

template <typename OutT, int ITEMS, int MINB>
cudaError_t launch_convert_v(const LaunchCtx &c, const char *in, size_t len, OutT *out, void *res, size_t tiles) {
  static int per_sm = 0;  // resident CTAs per SM for this instantiation (queried once)
  if (per_sm == 0) {
    int n = 0;
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k_convert_utf8<OutT, ITEMS, true, MINB>, kBlock, 0);
    if (e != cudaSuccess) return e;
    per_sm = n < 1 ? 1 : n;
  }
  const size_t cap = (size_t)c.sm_count * per_sm;
  const unsigned grid = (unsigned)(tiles < cap ? tiles : cap);
  k_convert_utf8<OutT, ITEMS, true, MINB><<<grid, kBlock, 0, c.stream>>>(in, len, out, c.scratch, c.desc, c.epoch,
                                                                        (uint32_t)tiles, static_cast<ResultPOD *>(res));
  count_launch(1);
  return cudaGetLastError();
}

// Resident CTAs per SM the kernels are compiled for (register budget = 65536 / (256 * MINB)).
// B200_TUNE_MINB=2|3|4 overrides the default for experiments (tools/, profiles/).
inline int tuned_minb(int dflt) {
  static int v = -1;
  if (v < 0) {
    const char *e = getenv("B200_TUNE_MINB");
    v = (e && e[0] >= '2' && e[0] <= '4' && e[1] == 0) ? e[0] - '0' : 0;
  }
  return v ? v : dflt;
}

template <typename OutT, int ITEMS>
cudaError_t launch_convert(const LaunchCtx &c, const char *in, size_t len, OutT *out, void *res) {
  const size_t tiles = tiles_for(in, len, ITEMS);
  if (tiles > c.desc_capacity || tiles > 0xFFFFFFF0ull) return cudaErrorInvalidValue;
  switch (tuned_minb(4)) {
    case 2: return launch_convert_v<OutT, ITEMS, 2>(c, in, len, out, res, tiles);
    case 3: return launch_convert_v<OutT, ITEMS, 3>(c, in, len, out, res, tiles);
    default: return launch_convert_v<OutT, ITEMS, 4>(c, in, len, out, res, tiles);
  }
}

}  // namespace

size_t utf8_to_utf32_tiles(const void *in, size_t len) { return tiles_for(in, len, kConv32Items); }

cudaError_t launch_write_result(void *res, int32_t error, unsigned long long count, cudaStream_t stream) {
  k_write_result<<<1, 1, 0, stream>>>(static_cast<ResultPOD *>(res), error, count);
  count_launch(1);
  return cudaGetLastError();
}
cudaError_t launch_write_full_result(void *res, int32_t error, unsigned long long in_count,
                                     unsigned long long out_count, cudaStream_t stream) {
  k_write_full_result<<<1, 1, 0, stream>>>(static_cast<FullResultPOD *>(res), error, in_count, out_count);
  count_launch(1);
  return cudaGetLastError();
}
cudaError_t launch_write_u64(unsigned long long *dst, unsigned long long v, cudaStream_t stream) {
  k_write_u64<<<1, 1, 0, stream>>>(dst, v);
  count_launch(1);
  return cudaGetLastError();
}
cudaError_t launch_scratch_init(Scratch *scr, cudaStream_t stream) {
  k_scratch_init<<<1, 1, 0, stream>>>(scr);
  count_launch(1);
  return cudaGetLastError();
}

cudaError_t launch_validate_utf8(const LaunchCtx &c, const char *in, size_t len, void *res) {
  const unsigned grid = reduction_grid(c, len, kStreamItems);
  k_validate_utf8<kStreamItems><<<grid, kBlock, 0, c.stream>>>(in, len, c.scratch, static_cast<ResultPOD *>(res));
  count_launch(1);
  return cudaGetLastError();
}

cudaError_t launch_count_utf8(const LaunchCtx &c, const char *in, size_t len, unsigned long long *count, int mode) {
  const unsigned grid = reduction_grid(c, len, kStreamItems);
  if (mode == 0) k_count_utf8<0, kStreamItems><<<grid, kBlock, 0, c.stream>>>(in, len, c.scratch, count);
  else k_count_utf8<1, kStreamItems><<<grid, kBlock, 0, c.stream>>>(in, len, c.scratch, count);
  count_launch(1);
  return cudaGetLastError();
}

cudaError_t launch_convert_utf8_to_utf32(const LaunchCtx &c, const char *in, size_t len, uint32_t *out, void *res) {
  return launch_convert<uint32_t, kConv32Items>(c, in, len, out, res);
}

}  // namespace b200
