// k_base64.cu — K7: WHATWG forgiving-base64 decode on sm_100a
//   implementation::base64_to_binary[_details] for `char` input, every base64_options /
//   last_chunk_handling_options value (reference src/generic/base64.h:40-246, src/scalar/base64.h:33-216,
//   src/tables/base64_tables.h:791-849).
//
// One launch, one pass over HBM:
//   1. each thread classifies its 16-byte granules through a 256-entry class table in shared memory
//      (built by the CTA from swar.h:b64_class — 0..63 sextet, 64 whitespace, 255 invalid);
//   2. valid characters are counted, block-scanned, and their sextets compacted into shared memory
//      (this is where whitespace disappears);
//   3. the tile's sextet count goes through the decoupled look-back scan, which also carries the last
//      sextet seen so far (descriptor `aux`), so a 4-sextet quantum may straddle any number of tiles;
//   4. output byte b of the stream is made of sextets r = 4*(b/3) + b%3 and r+1; the tile that owns sextet
//      r+1 writes it — with 16-byte coalesced stores through the same staging scheme as the transcoders;
//   5. the first invalid character is an atomicMin on its index; the CTA that finishes last strips the
//      trailing whitespace / '=' (block-cooperative backward scan), fetches the last <= 3 sextet
//      characters and applies the last-chunk and padding rules (swar.h:b64_finish).
#include "device_common.cuh"
#include "launch.h"

namespace b200 {

namespace {

constexpr int kItems = 4;                          // 16 KiB of text per tile
constexpr uint32_t kTileChars = kBlock * kItems * 16;

struct B64Smem {
  alignas(16) uint8_t sx[kTileChars + 16];         // sx[0] = carry-in sextet, sx[1+k] = k-th sextet of the tile
  alignas(16) uint8_t out[kTileChars / 4 * 3 + 48];
  uint8_t lut[256];
  uint32_t warp_tot[kWarps];
  uint32_t tile;
  uint32_t carry;
  unsigned long long excl;
  unsigned long long found;  // 1 + index, 0 = none
  int is_last;
};

__device__ __forceinline__ InView make_view(const void *p, size_t len_bytes) {
  InView v;
  const uintptr_t a = reinterpret_cast<uintptr_t>(p);
  v.base = reinterpret_cast<const uint4 *>(a & ~uintptr_t(15));
  v.vbeg = a & 15u;
  v.vend = v.vbeg + len_bytes;
  return v;
}

// Largest j in [0, end) with (lut[p[j]] == 64) != want_ws ... generalised: largest j whose class satisfies
// `sextet_only ? class <= 63 : class != 64`; -1 if none.  Whole CTA participates.
__device__ long long block_find_last(const uint8_t *p, long long end, bool sextet_only, B64Smem &sm) {
  constexpr long long kPer = 16;
  for (long long hi = end; hi > 0; hi -= kBlock * kPer) {
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

__global__ void __launch_bounds__(kBlock) k_base64_decode(const char *ptr, size_t len, uint8_t *out, Scratch *scr,
                                                          unsigned long long *desc, uint32_t epoch,
                                                          uint32_t num_tiles, uint32_t opt_url, uint32_t opt_both,
                                                          uint32_t opt_garbage, unsigned long long last_chunk,
                                                          FullResultPOD *res) {
  __shared__ B64Smem sm;
  const InView in = make_view(ptr, len);
  const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  sm.lut[threadIdx.x] = (uint8_t)b64_class(threadIdx.x, opt_url != 0, opt_both != 0);
  __syncthreads();

  while (true) {
    if (threadIdx.x == 0) sm.tile = atomicAdd(&scr->ticket, 1u);
    __syncthreads();
    const uint32_t tile = sm.tile;
    if (tile >= num_tiles) break;
    const unsigned long long g0 = ((unsigned long long)tile * kWarps + warp) * (32ull * kItems);

    uint32_t w[kItems][4];
    bool inside[kItems];
#pragma unroll
    for (int j = 0; j < kItems; j++) load_granule(in, g0 + (unsigned long long)j * 32u + lane, w[j], inside[j]);

    // ---- classify: per granule a 16-bit "is sextet" mask, the sextet values overwrite w ---------
    uint32_t vmask[kItems], cnt[kItems], off[kItems];
#pragma unroll
    for (int j = 0; j < kItems; j++) {
      const unsigned long long g = g0 + (unsigned long long)j * 32u + lane;
      uint32_t valid = 0, badm = 0;
#pragma unroll
      for (int k = 0; k < 4; k++) {
        uint32_t cls = 0;
#pragma unroll
        for (int b = 0; b < 4; b++) cls |= (uint32_t)sm.lut[(w[j][k] >> (8 * b)) & 0xFFu] << (8 * b);
        w[j][k] = cls;
        // class <= 63 <=> bits 6,7 clear ; class == 255 <=> bit 7 set
        const uint32_t is_sx = ~(cls | (cls << 1)) & kH;
        valid |= mask4(is_sx) << (4 * k);
        badm |= mask4(cls & kH) << (4 * k);
      }
      if (!inside[j]) {
        uint32_t r = 0;
#pragma unroll
        for (int k = 0; k < 4; k++) r |= mask4(inrange_mask_word(in, g, k)) << (4 * k);
        valid &= r;
        badm &= r;
      }
      vmask[j] = valid;
      cnt[j] = (uint32_t)__popc(valid);
      if (badm && !opt_garbage) {
        const unsigned long long pos = g * 16ull + (unsigned)(__ffs((int)badm) - 1) - in.vbeg;
        const unsigned long long key = err_key(pos, kInvalidBase64Character);
        if (key < ld_relaxed_u64(&scr->err_key)) report_error(scr, key);
      }
    }
    const uint32_t tile_total = block_exclusive_offsets<kItems>(cnt, off, sm.warp_tot);

    // ---- compact the sextets of this tile -------------------------------------------------------
#pragma unroll
    for (int j = 0; j < kItems; j++) {
      uint32_t o = 1u + off[j];
#pragma unroll
      for (int p = 0; p < 16; p++) {
        if ((vmask[j] >> p) & 1u) sm.sx[o++] = (uint8_t)((w[j][p >> 2] >> (8 * (p & 3))) & 0x3Fu);
      }
    }
    __syncthreads();

    // ---- look-back: global rank of the tile's first sextet + the sextet just before it ----------
    if (warp == 0) {
      unsigned long long excl;
      uint32_t aux;
      const uint32_t my_last = tile_total ? sm.sx[tile_total] : 0u;
      tile_lookback(desc, epoch, tile, tile_total, my_last, excl, aux);
      if (lane == 0) {
        sm.excl = excl;
        sm.sx[0] = (uint8_t)aux;
        if (tile == num_tiles - 1) st_relaxed_u64(&scr->acc0, excl + tile_total);
      }
    }
    __syncthreads();
    const unsigned long long R0 = sm.excl;
    const unsigned long long B0 = b64_bytes_from_sextets(R0);
    const uint32_t nbytes = (uint32_t)(b64_bytes_from_sextets(R0 + tile_total) - B0);
    uint8_t *gdst = out + B0;
    const uint32_t shift = staging_shift(gdst);
    const uint32_t m0 = (uint32_t)(R0 & 3ull);
    const uint32_t b0m = m0 ? m0 - 1u : 0u;

    // ---- pack: output byte t of this tile = sextets (k, k+1), k = 4*((b0m+t)/3) + (b0m+t)%3 - m0 ----
    for (uint32_t t = threadIdx.x; t < nbytes; t += kBlock) {
      const uint32_t a = b0m + t;
      const uint32_t q = a / 3u, m = a - 3u * q;
      const uint32_t k1 = 4u * q + m + 1u - m0;  // index into sx of the first sextet (sx[0] is rank R0-1)
      const uint32_t s0 = sm.sx[k1], s1 = sm.sx[k1 + 1u];
      sm.out[shift + t] = (uint8_t)((s0 << (2u + 2u * m)) | (s1 >> (4u - 2u * m)));
    }
    __syncthreads();
    copy_out_aligned<uint8_t>(sm.out, gdst, shift, nbytes);
    __syncthreads();
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
    srclen = (unsigned long long)(block_find_last(p, (long long)srclen, false, sm) + 1);
    equallocation = srclen;
    if (srclen > 0 && p[srclen - 1] == '=') {
      equallocation = srclen - 1;
      srclen--;
      equalsigns = 1;
      srclen = (unsigned long long)(block_find_last(p, (long long)srclen, false, sm) + 1);
      if (srclen > 0 && p[srclen - 1] == '=') {
        equallocation = srclen - 1;
        srclen--;
        equalsigns = 2;
      }
    }
  }
  const unsigned long long key = ld_relaxed_u64(&scr->err_key);
  const unsigned long long V = ld_relaxed_u64(&scr->acc0);
  const bool invalid = key != kNoError && (key >> 8) < srclen;
  uint32_t tail_val[3] = {0, 0, 0};
  uint64_t tail_pos[3] = {0, 0, 0};
  if (!invalid && srclen > 0) {
    long long end = (long long)srclen;
    const uint32_t idx = (uint32_t)(V & 3ull);
    for (uint32_t k = 0; k < idx; k++) {
      const long long f = block_find_last(p, end, true, sm);
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
      int error;
      uint64_t in_count, out_count;
      b64_finish(srclen, equalsigns, equallocation, V, opt_garbage != 0, last_chunk, tail_val, tail_pos, &error,
                 &in_count, &out_count);
      res->error = error;
      res->reserved_ = 0;
      res->input_count = in_count;
      res->output_count = out_count;
    }
    scratch_reset(scr);
  }
}

inline size_t tiles_for(const void *in, size_t len_bytes) {
  const size_t span = (reinterpret_cast<uintptr_t>(in) & 15u) + len_bytes;
  const size_t gran = (span + 15) / 16;
  const size_t per_tile = (size_t)kBlock * kItems;
  return (gran + per_tile - 1) / per_tile;
}

}  // namespace

size_t base64_tiles(const void *in, size_t len) { return tiles_for(in, len); }

cudaError_t launch_base64_to_binary(const LaunchCtx &c, const char *in, size_t len, char *out, uint64_t options,
                                    uint64_t last_chunk, void *full_res) {
  const size_t tiles = tiles_for(in, len);
  if (tiles > c.desc_capacity || tiles > 0xFFFFFFF0ull) return cudaErrorInvalidValue;
  int per_sm = 0;
  cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_base64_decode, kBlock, 0);
  if (e != cudaSuccess) return e;
  if (per_sm < 1) per_sm = 1;
  const size_t cap = (size_t)c.sm_count * per_sm;
  const unsigned grid = (unsigned)(tiles < cap ? tiles : cap);
  // reference include/simdutf/implementation.h:2782-2800 and src/scalar/base64.h:66-69
  const uint32_t url = (options & 1u) ? 1u : 0u;
  const uint32_t both = (options & 8u) ? 1u : 0u;
  const uint32_t garbage = (options == 4u || options == 5u || options == 12u) ? 1u : 0u;
  k_base64_decode<<<grid, kBlock, 0, c.stream>>>(in, len, reinterpret_cast<uint8_t *>(out), c.scratch, c.desc, c.epoch,
                                                 (uint32_t)tiles, url, both, garbage, last_chunk,
                                                 static_cast<FullResultPOD *>(full_res));
  count_launch(1);
  return cudaGetLastError();
}

}  // namespace b200
