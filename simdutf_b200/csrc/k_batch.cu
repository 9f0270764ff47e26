// k_batch.cu — sm_100a kernels for MANY SMALL STRINGS PER LAUNCH ("callers at the right
// granularity"; the reference's own callers of this shape are tools/sutf.cpp:131-336 and the per-string loops of its
// test suite, e.g. tests/validate_utf8_with_errors_tests.cpp:54-69).
//
// A host-pointer call costs a launch and a synchronisation (~25 us) however short its string; a batch pays that once.
// String i is data[offsets[i] .. offsets[i + 1]): one packed buffer plus n + 1 offsets, the layout of an Arrow string
// column.  ONE WARP PER STRING (grid-stride over the strings): a warp walks its string in 1 KiB chunks, every lane one
// 32-byte block in bit-plane form (bitplane.h), exactly the per-block logic of K1 / K2 / K3 — but the strings are
// independent, so there is no scan across warps, no scratch line and no finalisation: a warp writes its string's
// result itself.  Semantics per string are those of the single-string entry points:
//   validate_utf8_with_errors   reference src/scalar/utf8.h:102-200     -> result{error, count}
//   utf16_length_from_utf8      reference src/scalar/utf8.h:243-255     -> uint64
//   count_utf8                  reference src/scalar/utf8.h:230-241     -> uint64
//   convert_utf8_to_utf16le     reference src/scalar/utf8_to_utf16/utf8_to_utf16.h:128-255 -> result + units at
//                               out[out_offsets[i] ...] (out_offsets == nullptr: at out[offsets[i] ...]; a string
//                               never yields more units than it has bytes, so that layout needs no length pass)
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

// First error in byte positions [lo, hi) (virtual, clipped to the buffer) as position << 8 | code, or kNoError.
__device__ __noinline__ unsigned long long u8_first_error(const uint4 *base, unsigned long long vbeg,
                                                          unsigned long long vend, long long lo, long long hi) {
  if (lo < (long long)vbeg) lo = (long long)vbeg;
  if (hi > (long long)vend) hi = (long long)vend;
  const uint8_t *p = reinterpret_cast<const uint8_t *>(base) + vbeg;
  const unsigned long long len = vend - vbeg;
  auto at = [p](unsigned long long j) -> uint32_t { return (uint32_t)p[j]; };
  for (long long v = lo; v < hi; v++) {
    const unsigned long long i = (unsigned long long)v - vbeg;
    const int code = u8_verdict(at, i, len);
    if (code != kSuccess) return err_key(i, code);
  }
  return kNoError;
}

__device__ __forceinline__ bool tail_truncated(const InView &in) {
  const long long e = (long long)in.vend;
  const uint8_t *p = reinterpret_cast<const uint8_t *>(in.base);
  const uint32_t b1 = p[e - 1];
  const uint32_t b2 = (e - 2 >= (long long)in.vbeg) ? p[e - 2] : 0u;
  const uint32_t b3 = (e - 3 >= (long long)in.vbeg) ? p[e - 3] : 0u;
  return u8_incomplete_tail(b1, b2, b3);
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

// The lane's 32-byte block at virtual offset r0 (zero outside the buffer), the word before it, the byte after it.
__device__ __forceinline__ void load_block(const InView &in, unsigned long long r0, uint32_t (&B)[8], uint32_t &pw,
                                           uint32_t &nbyte) {
  bool ins;
  load_granule(in, r0 >> 4, &B[0], ins);
  load_granule(in, (r0 >> 4) + 1ull, &B[4], ins);
  pw = load_word_guarded(in, (long long)(r0 >> 2) - 1);
  const unsigned long long np = r0 + 32ull;
  nbyte = (np >= in.vbeg && np < in.vend) ? (uint32_t)__ldg(reinterpret_cast<const uint8_t *>(in.base) + np) : 0u;
}

constexpr int kBatchThreads = 256;

// MODE 0 validate (res = ResultPOD[n]), 1 count_utf8, 2 utf16_length_from_utf8 (res = uint64[n])
template <int MODE>
__global__ void __launch_bounds__(kBatchThreads) k_utf8_batch(const char *data, const unsigned long long *offs,
                                                             unsigned long long n, void *res) {
  const unsigned lane = threadIdx.x & 31u;
  const unsigned long long warps = (unsigned long long)gridDim.x * (kBatchThreads / 32);
  for (unsigned long long s = (unsigned long long)blockIdx.x * (kBatchThreads / 32) + (threadIdx.x >> 5); s < n; s += warps) {
    const unsigned long long o0 = offs[s], len = offs[s + 1] - o0;
    const InView in = make_view(data + o0, len);
    unsigned long long key = kNoError, total = 0;
    for (unsigned long long c0 = 0; c0 < in.vend && len; c0 += 1024ull) {
      const unsigned long long r0 = c0 + 32ull * lane;
      uint32_t B[8], pw, nbyte;
      load_block(in, r0, B, pw, nbyte);
      const uint32_t inr = range_mask32(in, r0);
      if (MODE == 0) {
        uint32_t hi = pw;
#pragma unroll
        for (int i = 0; i < 8; i++) hi |= B[i];
        const bool last = c0 + 1024ull >= in.vend;
        if (!last && !__any_sync(kFull, (hi & kH) != 0u)) continue;
        bp::VCarry vc = bp::vcarry_from_word(pw);
        bp::transpose_in(B);
        bool bad = bp::utf8_check_block(B, vc) != 0u;
        if (r0 < in.vend && in.vend <= r0 + 32ull && tail_truncated(in)) bad = true;
        if (bad) key = u8_first_error(in.base, in.vbeg, in.vend, (long long)r0 - 3, (long long)r0 + 32);
        // every chunk before this one was clean, so the smallest key of this chunk is the string's first error
        if (__any_sync(kFull, key != kNoError)) break;
      } else {
        bp::transpose_in(B);
        const uint32_t nc = (~B[7] | B[6]) & inr;                           // bytes that start a character
        uint32_t m = (uint32_t)__popc(nc);
        if (MODE == 2) m += (uint32_t)__popc(B[7] & B[6] & B[5] & B[4] & inr);  // 4-byte leads: a second unit
        total += m;
      }
    }
    if (MODE == 0) {
      key = warp_min_u64(key);
      if (lane == 0) {
        ResultPOD *r = static_cast<ResultPOD *>(res) + s;
        r->error = key == kNoError ? kSuccess : (int32_t)(key & 0xFFu);
        r->reserved_ = 0;
        r->count = key == kNoError ? len : key >> 8;
      }
    } else {
      total = warp_sum_u64(total);
      if (lane == 0) static_cast<unsigned long long *>(res)[s] = total;
    }
  }
}

// convert_utf8_to_utf16le/be[_with_errors], one warp per string: the warp keeps its own running output offset, so the
// units go straight from the registers to global memory (2-byte stores; these strings are short by assumption).
template <bool BE>
__global__ void __launch_bounds__(kBatchThreads) k_utf8_to_utf16_batch(const char *data, const unsigned long long *offs,
                                                                      unsigned long long n, uint16_t *out,
                                                                      const unsigned long long *out_offs, ResultPOD *res) {
  const unsigned lane = threadIdx.x & 31u;
  const unsigned long long warps = (unsigned long long)gridDim.x * (kBatchThreads / 32);
  for (unsigned long long s = (unsigned long long)blockIdx.x * (kBatchThreads / 32) + (threadIdx.x >> 5); s < n; s += warps) {
    const unsigned long long o0 = offs[s], len = offs[s + 1] - o0;
    const InView in = make_view(data + o0, len);
    uint16_t *dst = out + (out_offs ? out_offs[s] : o0);
    unsigned long long key = kNoError, produced = 0;
    // a string that starts with a continuation byte is invalid at position 0 and emits nothing (bitplane.h)
    const bool poison = len && ((uint32_t)__ldg(reinterpret_cast<const uint8_t *>(in.base) + in.vbeg) & 0xC0u) == 0x80u;
    for (unsigned long long c0 = 0; c0 < in.vend && len; c0 += 1024ull) {
      const unsigned long long r0 = c0 + 32ull * lane;
      uint32_t B[8], pw, nbyte;
      load_block(in, r0, B, pw, nbyte);
      bp::Carry carry = bp::carry_from_word(pw);
      const uint32_t next_nc = ((nbyte & 0xC0u) != 0x80u) ? 1u : 0u;
      bp::transpose_in(B);
      uint32_t m = bp::emit16_mask(B, carry.l4, next_nc) & range_mask32(in, r0);
      if (poison) m = 0;
      const uint32_t cnt = (uint32_t)__popc(m);
      uint32_t incl = cnt;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(kFull, incl, o);
        if (lane >= (unsigned)o) incl += t;
      }
      const uint32_t wtot = __shfl_sync(kFull, incl, 31);
      uint32_t U[16];
      bool bad = bp::utf8_to_utf16_block<true>(B, carry, U) != 0u;
      if (r0 < in.vend && in.vend <= r0 + 32ull && tail_truncated(in)) bad = true;
      if (bad) key = u8_first_error(in.base, in.vbeg, in.vend, (long long)r0 - 3, (long long)r0 + 32);
      if (BE) {
#pragma unroll
        for (int k = 0; k < 8; k++) {
          const uint32_t t = U[k];
          U[k] = U[k + 8];
          U[k + 8] = t;
        }
      }
      bp::transpose_out16(U);
      uint16_t *q = dst + produced + (incl - cnt);
#pragma unroll
      for (int i = 0; i < 16; i++) {
        if (m & (1u << i)) *q++ = (uint16_t)U[i];
      }
#pragma unroll
      for (int i = 0; i < 16; i++) {
        if (m & (1u << (16 + i))) *q++ = (uint16_t)(U[i] >> 16);
      }
      produced += wtot;
      if (__any_sync(kFull, key != kNoError)) break;
    }
    key = warp_min_u64(key);
    if (lane == 0) {
      res[s].error = key == kNoError ? kSuccess : (int32_t)(key & 0xFFu);
      res[s].reserved_ = 0;
      res[s].count = key == kNoError ? produced : key >> 8;
    }
  }
}

inline unsigned batch_grid(int sm_count, unsigned long long n) {
  const unsigned long long ctas = (n + kBatchThreads / 32 - 1) / (kBatchThreads / 32);
  const unsigned long long cap = (unsigned long long)sm_count * 8;
  return (unsigned)(ctas < 1 ? 1 : (ctas < cap ? ctas : cap));
}

}  // namespace

cudaError_t launch_utf8_batch(int sm_count, cudaStream_t stream, int mode, const char *data, const unsigned long long *offs,
                              unsigned long long n, void *res) {
  if (n == 0) return cudaSuccess;
  const unsigned grid = batch_grid(sm_count, n);
  if (mode == 0) k_utf8_batch<0><<<grid, kBatchThreads, 0, stream>>>(data, offs, n, res);
  else if (mode == 1) k_utf8_batch<1><<<grid, kBatchThreads, 0, stream>>>(data, offs, n, res);
  else k_utf8_batch<2><<<grid, kBatchThreads, 0, stream>>>(data, offs, n, res);
  count_launch(1);
  return cudaGetLastError();
}

cudaError_t launch_utf8_to_utf16_batch(int sm_count, cudaStream_t stream, bool big_endian, const char *data,
                                       const unsigned long long *offs, unsigned long long n, uint16_t *out,
                                       const unsigned long long *out_offs, void *res) {
  if (n == 0) return cudaSuccess;
  const unsigned grid = batch_grid(sm_count, n);
  if (big_endian)
    k_utf8_to_utf16_batch<true><<<grid, kBatchThreads, 0, stream>>>(data, offs, n, out, out_offs, static_cast<ResultPOD *>(res));
  else
    k_utf8_to_utf16_batch<false><<<grid, kBatchThreads, 0, stream>>>(data, offs, n, out, out_offs, static_cast<ResultPOD *>(res));
  count_launch(1);
  return cudaGetLastError();
}

}  // namespace b200
