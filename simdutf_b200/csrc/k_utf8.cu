// k_utf8.cu — sm_100a reduction kernels whose input is UTF-8:
//   K1  validate_utf8_with_errors           (reference src/scalar/utf8.h:102-200)
//   K2  count_utf8 / utf16_length_from_utf8 (reference src/scalar/utf8.h:230-255)
// (the transcoders UTF-8 -> UTF-16LE / UTF-32 live in k_utf8_to_utf16.cu)
//
// Data layout: the input is read once, as 16-byte granules with fully coalesced 128-bit streaming loads
// (lane l of a warp owns granule g0 + j*32 + l, j = 0..ITEMS-1, so one load instruction covers 512
// contiguous bytes).  Counting is a popcount reduction.  Validation skips all-ASCII 2 KiB chunks after one
// OR-reduction and a vote; any other chunk is re-laid-out through shared memory so that every lane holds 64
// contiguous bytes and checked in bit-plane form (bitplane.h); first error = atomicMin of position<<8|code.
#include <cstdlib>
#include <type_traits>

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
    // (Measured and dropped: pulling the warp's next chunk into L2 with prefetch.global.L2 — ncu puts 27 % of the stall
    // samples of the non-ASCII path on the wait for the chunk's own loads — made mixed text 5 % and ASCII 5-15 % SLOWER;
    // 3, 5 or 6 resident CTAs per SM (70 / 48 / 40 registers) are within 2 % of each other.)
    // interior chunks (every granule inside the buffer; all but the first and the last chunk) skip the per-granule
    // range tests: the mixed-text path is ALU-bound (ncu: ALU pipe 79 %), and the guards were a tenth of its instructions
    if (g0 * 16ull >= in.vbeg && (g0 + chunk_gran) * 16ull <= in.vend) {
#pragma unroll
      for (int j = 0; j < ITEMS; j++) {
        const uint4 v = ldg_stream_v4(in.base + g0 + (unsigned long long)j * 32u + lane);
        w[j][0] = v.x; w[j][1] = v.y; w[j][2] = v.z; w[j][3] = v.w;
      }
    } else {
      bool inside;
#pragma unroll
      for (int j = 0; j < ITEMS; j++) load_granule(in, g0 + (unsigned long long)j * 32u + lane, w[j], inside);
    }
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
    if (__any_sync(kFull, badblocks != 0u)) {
      // An error that is already on record in front of this chunk makes the chunk — and every later chunk of this
      // warp — irrelevant (the first error wins, the key only ever decreases): stop reading.  Keeps validation of
      // text in the wrong encoding (detect_encodings) from locating an error in every block.
      unsigned long long cur = ld_relaxed_u64(&scr->err_key);
      cur = __shfl_sync(kFull, cur, 0);
      const long long first = (long long)(g0 * 16ull) - 3 - (long long)in.vbeg;
      if (cur != kNoError && (long long)(cur >> 8) < first) break;
      // The flagged lanes take turns, lowest first, and the warp stops at the first one that pins an error down: every
      // error lies in the window of a flagged block, windows grow with the lane, so the first hit is the chunk's first
      // error.  (All flagged lanes searching at once cost text in the wrong encoding one atomic per LANE before the
      // first error was on record: validate_utf8 on 1 GiB of UTF-16 text took 0.14 ms to give up.)
      unsigned bm = __ballot_sync(kFull, badblocks != 0u);
      while (bm) {
        const unsigned l = (unsigned)__ffs((int)bm) - 1u;
        int found = 0;
        if (lane == l) {
#pragma unroll
          for (int j = 0; j < ITEMS / 2; j++) {
            const long long b0 = (long long)(r0 + 32ull * j);
            if (!found && (badblocks & (1u << j))) found = u8_locate_error(in, scr, b0 - 3, b0 + 32) ? 1 : 0;
          }
        }
        if (__shfl_sync(kFull, found, l)) break;
        bm &= bm - 1u;
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
    // Chunks wholly inside the buffer (all but the first and the last) take unguarded loads and none of the per-granule
    // range tests: the same split as in k_validate_utf8, where the guards were a tenth of the instructions.
    auto chunk_count = [&](auto interior_tag) -> uint32_t {
      constexpr bool kInterior = decltype(interior_tag)::value;
      uint32_t w[ITEMS][4];
      bool inside[ITEMS];
#pragma unroll
      for (int j = 0; j < ITEMS; j++) {
        if constexpr (kInterior) {
          const uint4 v = ldg_stream_v4(in.base + g0 + (unsigned long long)j * 32u + lane);
          w[j][0] = v.x; w[j][1] = v.y; w[j][2] = v.z; w[j][3] = v.w;
          inside[j] = true;
        } else {
          load_granule(in, g0 + (unsigned long long)j * 32u + lane, w[j], inside[j]);
        }
      }
      uint32_t cnt = 0;
#pragma unroll
      for (int j = 0; j < ITEMS; j++) {
        const unsigned long long g = g0 + (unsigned long long)j * 32u + lane;
        if ((kInterior || inside[j]) && MODE == 1) {
          // utf16_length: only the HIGH NIBBLE of a byte matters (a byte starts a character unless it is 10xx, and brings
          // a second unit when it is 1111), so two words are packed into one word of eight high nibbles and classified
          // together: per nibble n3 n2 n1 n0, "starts a character" = ~n3 | n2 lands on bit 3, ">= 0xF0" = n3 n2 n1 n0 on
          // bit 2, and the second pair of words goes to bits 1 and 0: ONE popcount per granule, half the logic
          // instructions of the per-word form (ncu: this kernel was ALU-bound at 72 % of the pipe and 0.88 of the copy
          // bandwidth).
          uint32_t m = 0;
#pragma unroll
          for (int k = 0; k < 4; k += 2) {
            const uint32_t x = bp::bitsel(w[j][k], w[j][k + 1] >> 4, 0xF0F0F0F0u);
            const uint32_t s1 = x << 1;
            const uint32_t nc = (~x | s1) & 0x88888888u;
            const uint32_t a = x & s1;                       // bit 3: n3 n2, bit 1: n1 n0
            const uint32_t f0 = a & (a << 2) & 0x88888888u;  // bit 3: n3 n2 n1 n0
            const uint32_t mk = nc | (f0 >> 1);
            m |= k == 0 ? mk : (mk >> 2);
          }
          cnt += (uint32_t)__popc(m);
        } else if (kInterior || inside[j]) {
          // the masks of the four words share one popcount: word k's bit 7s move down by (3 - k)
          uint32_t m = 0;
#pragma unroll
          for (int k = 0; k < 4; k++) m |= u8_noncont(w[j][k]) >> (3 - k);
          cnt += (uint32_t)__popc(m);
        } else {
#pragma unroll
          for (int k = 0; k < 4; k++) {
            uint32_t m = u8_noncont(w[j][k]);
            if (MODE == 1) m |= u8_ge_f0(w[j][k]) >> 1;  // bit 6 of the same byte: one popc counts both
            const uint32_t r = inrange_mask_word(in, g, k);
            m &= r | (r >> 1);
            cnt += (uint32_t)__popc(m);
          }
        }
      }
      return cnt;
    };
    // (utf16_length only: 0.895 -> 0.955 of the copy bandwidth inside the bench step; count_utf8 — one popcount per
    // granule, at 1.0 with the guarded loads — lost 5 % with the split and keeps the one path)
    const bool interior = MODE == 1 && g0 * 16ull >= in.vbeg && (g0 + chunk_gran) * 16ull <= in.vend;  // warp-uniform
    const uint32_t cnt = interior ? chunk_count(std::true_type{}) : chunk_count(std::false_type{});
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

inline unsigned reduction_grid(const LaunchCtx &c, size_t len_bytes, int items) {
  const unsigned long long chunks = (len_bytes + 16 + 511ull * items) / (512ull * items);
  const unsigned long long ctas = (chunks + kWarps - 1) / kWarps;
  const unsigned long long cap = (unsigned long long)c.sm_count * 8;  // 8 CTAs x 256 threads = full SM
  return (unsigned)(ctas < 1 ? 1 : (ctas < cap ? ctas : cap));
}

}  // namespace


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


}  // namespace b200
