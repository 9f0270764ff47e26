// k_utf16.cu — sm_100a kernels whose input is UTF-16LE:
//   K5  count_utf16le / utf8_length_from_utf16le   (reference src/scalar/utf16.h:69-94)
//       validate_utf16le_with_errors                (reference src/scalar/utf16.h:39-67)
// (convert_utf16le_to_utf8 lives in k_utf16_to_utf8.cu)
//
// Same data path as k_utf8.cu: 16-byte granules (8 units) loaded with coalesced 128-bit streaming loads,
// neighbour unit by warp shuffle.  Surrogate verdicts are exact per unit, so the first error is
// a plain atomicMin of (unit index << 8 | SURROGATE).
#include <type_traits>

#include "device_common.cuh"
#include "launch.h"

namespace b200 {

namespace {

__device__ __forceinline__ InView make_view16(const uint16_t *p, size_t len_units) {
  InView v;
  const uintptr_t a = reinterpret_cast<uintptr_t>(p);
  v.base = reinterpret_cast<const uint4 *>(a & ~uintptr_t(15));
  v.vbeg = a & 15u;
  v.vend = v.vbeg + 2ull * len_units;
  return v;
}

// 8-bit mask of the units of granule g that lie inside the buffer.
__device__ __forceinline__ uint32_t inrange_units(const InView &in, unsigned long long g) {
  uint32_t m = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) {
    const unsigned long long pos = g * 16ull + 2u * i;
    if (pos >= in.vbeg && pos < in.vend) m |= 1u << i;
  }
  return m;
}

// Reports the first bad surrogate of one granule, if any.  Zero filler outside the buffer is neither a
// high nor a low surrogate, so buffer edges need no special case.
__device__ __forceinline__ void check_surrogates(const InView &in, Scratch *scr, unsigned long long g,
                                                 const uint32_t w[4], uint32_t pw, uint32_t nw, uint32_t valid) {
  // any unit in D800..DFFF?  (u & 0xF800) == 0xD800, on both halves of each word at once
  uint32_t any = 0;
#pragma unroll
  for (int k = 0; k < 4; k++) any |= u16_sur_flags(w[k]);  // swar.h; the screen below reuses these flags
  if (!any) return;
  // Surrogates present (common in real text: every supplementary character is a pair).  Exact screen on both halves
  // of each word at once: the low surrogates must be exactly the units behind the high surrogates — the unit before
  // the granule (upper half of pw) and the one after it (lower half of nw) included.  Only a granule that fails it
  // is searched unit by unit.
  if (!u16_pairing_screen(w, pw, nw)) return;  // swar.h (host-tested)
#pragma unroll
  for (int i = 0; i < 8; i++) {
    if (!((valid >> i) & 1u)) continue;
    const uint32_t u = u16_unit(w, i);
    const uint32_t pu = i == 0 ? (pw >> 16) : u16_unit(w, i - 1);
    const uint32_t nu = i == 7 ? (nw & 0xFFFFu) : u16_unit(w, i + 1);
    if (u16_bad(u, pu, true, nu, true)) {
      const unsigned long long idx = (g * 16ull + 2u * i - in.vbeg) >> 1;
      const unsigned long long cur = ld_relaxed_u64(&scr->err_key);
      const unsigned long long key = err_key(idx, kSurrogate);
      if (key < cur) report_error(scr, key);
      return;
    }
  }
}

__device__ __forceinline__ void write_result_from_key16(ResultPOD *res, unsigned long long key,
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
// K5: MODE 0 count_utf16le, MODE 1 utf8_length_from_utf16le, MODE 2 validate_utf16le_with_errors.
// ---------------------------------------------------------------------------------------------
// BE = true: the units are big-endian (UTF-16BE twins, reference include/simdutf/implementation.h:3443-3569,
// 4299-4368, 4783-4800): every loaded word is byte-swapped within its halves (one PRMT) and the rest is unchanged.
__device__ __forceinline__ uint32_t swap16x2(uint32_t w) { return __byte_perm(w, 0u, 0x2301); }

template <int MODE, int ITEMS, bool BE>
__global__ void __launch_bounds__(kBlock) k_scan_utf16(const uint16_t *ptr, size_t len, Scratch *scr, void *out) {
  __shared__ unsigned long long s_part[kWarps];
  const InView in = make_view16(ptr, len);
  const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  const unsigned long long ngran = (in.vend + 15ull) >> 4;
  const unsigned long long chunk_gran = 32ull * ITEMS;
  const unsigned long long nchunks = (ngran + chunk_gran - 1) / chunk_gran;
  const unsigned long long nwarps = (unsigned long long)gridDim.x * kWarps;
  unsigned long long total = 0;
  for (unsigned long long chunk = (unsigned long long)blockIdx.x * kWarps + warp; chunk < nchunks; chunk += nwarps) {
    const unsigned long long g0 = chunk * chunk_gran;
    // chunks wholly inside the buffer (all but the first and the last): unguarded loads, no per-granule range tests
    const bool interior = g0 * 16ull >= in.vbeg && (g0 + chunk_gran) * 16ull <= in.vend;  // warp-uniform
    uint32_t w[ITEMS][4];
    bool inside[ITEMS];
    if (interior) {
#pragma unroll
      for (int j = 0; j < ITEMS; j++) {
        const uint4 v = ldg_stream_v4(in.base + g0 + (unsigned long long)j * 32u + lane);
        w[j][0] = v.x; w[j][1] = v.y; w[j][2] = v.z; w[j][3] = v.w;
        inside[j] = true;
      }
    } else {
#pragma unroll
      for (int j = 0; j < ITEMS; j++) load_granule(in, g0 + (unsigned long long)j * 32u + lane, w[j], inside[j]);
    }
    if (BE && MODE != 2) {
#pragma unroll
      for (int j = 0; j < ITEMS; j++) {
#pragma unroll
        for (int k = 0; k < 4; k++) w[j][k] = swap16x2(w[j][k]);
      }
    }
    if (MODE == 2) {
      uint32_t pw[ITEMS], nw[ITEMS];
      neighbour_words<ITEMS>(in, g0, w, pw, nw);
      if (BE) {
#pragma unroll
        for (int j = 0; j < ITEMS; j++) {
#pragma unroll
          for (int k = 0; k < 4; k++) w[j][k] = swap16x2(w[j][k]);
          pw[j] = swap16x2(pw[j]);
          nw[j] = swap16x2(nw[j]);
        }
      }
#pragma unroll
      for (int j = 0; j < ITEMS; j++) {
        const unsigned long long g = g0 + (unsigned long long)j * 32u + lane;
        const uint32_t valid = inside[j] ? 0xFFu : inrange_units(in, g);
        check_surrogates(in, scr, g, w[j], pw[j], nw[j], valid);
      }
    } else {
      auto count_chunk = [&](auto interior_tag) -> uint32_t {
        constexpr bool kInterior = decltype(interior_tag)::value;
        uint32_t cnt = 0;
#pragma unroll
        for (int j = 0; j < ITEMS; j++) {
          const unsigned long long g = g0 + (unsigned long long)j * 32u + lane;
          if (kInterior || inside[j]) {
            // 16-bit-lane SWAR: bit 15 of a lane <=> the predicate holds for that unit; the four words of the granule
            // share one popcount (POPC issues at a quarter of the logic rate: one per word made the kernel POPC-bound)
            uint32_t m = 0;
#pragma unroll
            for (int k = 0; k < 4; k++) {
              const uint32_t x = w[j][k];
              if (MODE == 0) {
                const uint32_t z = (x ^ 0xDC00DC00u) & 0xFC00FC00u;         // zero lane <=> low surrogate
                const uint32_t notlow = (z >> 1) + 0x7E007E00u;
                m |= (notlow & 0x80008000u) >> k;
              } else {
                const uint32_t h = x >> 1;
                const uint32_t ge80 = (h & 0x7FC07FC0u) + 0x7FC07FC0u;
                const uint32_t ge800 = (h & 0x7C007C00u) + 0x7C007C00u;
                const uint32_t z = (x ^ 0xD800D800u) & 0xF800F800u;         // zero lane <=> surrogate
                const uint32_t notsur = (z >> 1) + 0x7C007C00u;
                m |= ((ge80 & 0x80008000u) | ((ge800 & notsur & 0x80008000u) >> 1)) >> (2 * k);
              }
            }
            cnt += (MODE == 0 ? 0u : 8u) + (uint32_t)__popc(m);
          } else {
            const uint32_t valid = inrange_units(in, g);
#pragma unroll
            for (int i = 0; i < 8; i++) {
              const uint32_t u = u16_unit(w[j], i);
              const uint32_t c = MODE == 0 ? (uint32_t)((u & 0xFC00u) != 0xDC00u) : u16_utf8_bytes(u);
              cnt += ((valid >> i) & 1u) ? c : 0u;
            }
          }
        }
        return cnt;
      };
      total += interior ? count_chunk(std::true_type{}) : count_chunk(std::false_type{});
    }
  }
  if (MODE != 2) {
    total = warp_sum_u64(total);
    if (lane == 0) s_part[warp] = total;
    __syncthreads();
    if (threadIdx.x == 0) {
      unsigned long long t = 0;
#pragma unroll
      for (int i = 0; i < kWarps; i++) t += s_part[i];
      if (t) atomicAdd(&scr->acc0, t);
    }
  }
  if (grid_last_thread(scr)) {
    if (MODE == 2) write_result_from_key16(static_cast<ResultPOD *>(out), ld_relaxed_u64(&scr->err_key), len);
    else *static_cast<unsigned long long *>(out) = ld_relaxed_u64(&scr->acc0);
    scratch_reset(scr);
  }
}

constexpr int kStreamItems = 4;

inline unsigned reduction_grid(const LaunchCtx &c, size_t len_bytes, int items) {
  const unsigned long long chunks = (len_bytes + 16 + 511ull * items) / (512ull * items);
  const unsigned long long ctas = (chunks + kWarps - 1) / kWarps;
  const unsigned long long cap = (unsigned long long)c.sm_count * 8;
  return (unsigned)(ctas < 1 ? 1 : (ctas < cap ? ctas : cap));
}

}  // namespace

cudaError_t launch_count_utf16(const LaunchCtx &c, const uint16_t *in, size_t len, unsigned long long *count, int mode,
                               bool big_endian) {
  const unsigned grid = reduction_grid(c, 2 * len, kStreamItems);
  if (mode == 0) {
    if (big_endian) k_scan_utf16<0, kStreamItems, true><<<grid, kBlock, 0, c.stream>>>(in, len, c.scratch, count);
    else k_scan_utf16<0, kStreamItems, false><<<grid, kBlock, 0, c.stream>>>(in, len, c.scratch, count);
  } else {
    if (big_endian) k_scan_utf16<1, kStreamItems, true><<<grid, kBlock, 0, c.stream>>>(in, len, c.scratch, count);
    else k_scan_utf16<1, kStreamItems, false><<<grid, kBlock, 0, c.stream>>>(in, len, c.scratch, count);
  }
  count_launch(1);
  return cudaGetLastError();
}

cudaError_t launch_validate_utf16(const LaunchCtx &c, const uint16_t *in, size_t len, void *res, bool big_endian) {
  const unsigned grid = reduction_grid(c, 2 * len, kStreamItems);
  if (big_endian) k_scan_utf16<2, kStreamItems, true><<<grid, kBlock, 0, c.stream>>>(in, len, c.scratch, res);
  else k_scan_utf16<2, kStreamItems, false><<<grid, kBlock, 0, c.stream>>>(in, len, c.scratch, res);
  count_launch(1);
  return cudaGetLastError();
}

// change_endianness_utf16 (reference include/simdutf/implementation.h:4567-4584): out[i] = byteswap(in[i]).
// Pure streaming: 16-byte vectors where both pointers allow it, units at the ragged ends.
__global__ void __launch_bounds__(kBlock) k_swap_utf16(const uint16_t *in, size_t len, uint16_t *out) {
  const size_t tid = (size_t)blockIdx.x * kBlock + threadIdx.x, nthreads = (size_t)gridDim.x * kBlock;
  const uintptr_t ai = reinterpret_cast<uintptr_t>(in), ao = reinterpret_cast<uintptr_t>(out);
  if (((ai ^ ao) & 15u) == 0) {  // same misalignment: a common 16-byte grid exists
    size_t head = ((16u - (ai & 15u)) & 15u) >> 1;
    if (head > len) head = len;
    const size_t nvec = (len - head) >> 3;
    const uint4 *vi = reinterpret_cast<const uint4 *>(in + head);
    uint4 *vo = reinterpret_cast<uint4 *>(out + head);
    for (size_t v = tid; v < nvec; v += nthreads) {
      uint4 x = ld_cs_v4(vi + v);
      x.x = swap16x2(x.x); x.y = swap16x2(x.y); x.z = swap16x2(x.z); x.w = swap16x2(x.w);
      stg_stream_v4(vo + v, x);
    }
    const size_t done = head + nvec * 8;
    for (size_t i = tid; i < head; i += nthreads) out[i] = (uint16_t)((in[i] >> 8) | (in[i] << 8));
    for (size_t i = done + tid; i < len; i += nthreads) out[i] = (uint16_t)((in[i] >> 8) | (in[i] << 8));
  } else {
    for (size_t i = tid; i < len; i += nthreads) out[i] = (uint16_t)((in[i] >> 8) | (in[i] << 8));
  }
}

// to_well_formed_utf16le/be (reference src/scalar/utf16.h:141-166; include/simdutf/implementation.h:3498-3531):
// out[i] = U+FFFD (in the buffer's byte order) where in[i] is a lone surrogate — a low one not behind a high one, a
// high one not in front of a low one or at the end — else in[i].  The rule is u16_bad() per unit with its two
// neighbours.  in == out is allowed: a neighbour that another thread has already replaced was a LONE surrogate, so it
// did not pair with this unit and U+FFFD in its place (not a surrogate) leads to the same decision.
template <bool BE>
__device__ __forceinline__ uint32_t well_formed_unit(uint32_t u, uint32_t pu, bool has_prev, uint32_t nu, bool has_next) {
  const uint32_t a = BE ? swap16x2(u) & 0xFFFFu : u, pa = BE ? swap16x2(pu) & 0xFFFFu : pu, na = BE ? swap16x2(nu) & 0xFFFFu : nu;
  return u16_bad(a, pa, has_prev, na, has_next) ? (BE ? 0xFDFFu : 0xFFFDu) : u;
}
template <bool BE>
__global__ void __launch_bounds__(kBlock) k_well_formed_utf16(const uint16_t *in, size_t len, uint16_t *out) {
  const size_t tid = (size_t)blockIdx.x * kBlock + threadIdx.x, nthreads = (size_t)gridDim.x * kBlock;
  const uintptr_t ai = reinterpret_cast<uintptr_t>(in), ao = reinterpret_cast<uintptr_t>(out);
  auto element = [&](size_t i) {
    const uint32_t pu = i ? in[i - 1] : 0u, nu = i + 1 < len ? in[i + 1] : 0u;
    out[i] = (uint16_t)well_formed_unit<BE>(in[i], pu, i > 0, nu, i + 1 < len);
  };
  if (((ai ^ ao) & 15u) == 0) {  // same misalignment: a common 16-byte grid exists
    size_t head = ((16u - (ai & 15u)) & 15u) >> 1;
    if (head > len) head = len;
    const size_t nvec = (len - head) >> 3;
    const uint4 *vi = reinterpret_cast<const uint4 *>(in + head);
    uint4 *vo = reinterpret_cast<uint4 *>(out + head);
    for (size_t v = tid; v < nvec; v += nthreads) {
      const size_t i0 = head + 8 * v;
      const uint4 x = ld_cs_v4(vi + v);
      const uint32_t w[4] = {x.x, x.y, x.z, x.w};
      uint32_t u[10];
      u[0] = i0 ? in[i0 - 1] : 0u;
      u[9] = i0 + 8 < len ? in[i0 + 8] : 0u;
#pragma unroll
      for (int k = 0; k < 8; k++) u[k + 1] = (w[k >> 1] >> (16 * (k & 1))) & 0xFFFFu;
      uint32_t o[4];
      // Exact screen (the one of check_surrogates): the low surrogates of this vector must be exactly the units
      // behind the high surrogates, the unit before and the unit after the vector included.  Well-formed vectors
      // — all of them in valid text — are copied through.
      uint32_t ws[4];
#pragma unroll
      for (int k = 0; k < 4; k++) ws[k] = BE ? swap16x2(w[k]) : w[k];
      const uint32_t wrong = u16_pairing_screen_tags(ws, BE ? swap16x2(u[0] << 16) : u[0] << 16, BE ? swap16x2(u[9]) : u[9]);
      if (!wrong) {
        o[0] = x.x; o[1] = x.y; o[2] = x.z; o[3] = x.w;
      } else {
#pragma unroll
        for (int k = 0; k < 8; k += 2) {
          const uint32_t lo = well_formed_unit<BE>(u[k + 1], u[k], i0 + k > 0, u[k + 2], i0 + k + 1 < len);
          const uint32_t hi = well_formed_unit<BE>(u[k + 2], u[k + 1], true, u[k + 3], i0 + k + 2 < len);
          o[k >> 1] = lo | (hi << 16);
        }
      }
      stg_stream_v4(vo + v, make_uint4(o[0], o[1], o[2], o[3]));
    }
    const size_t done = head + nvec * 8;
    for (size_t i = tid; i < head; i += nthreads) element(i);
    for (size_t i = done + tid; i < len; i += nthreads) element(i);
  } else {
    for (size_t i = tid; i < len; i += nthreads) element(i);
  }
}

cudaError_t launch_to_well_formed_utf16(const LaunchCtx &c, const uint16_t *in, size_t len, uint16_t *out, bool big_endian) {
  const unsigned long long want = (len / 8 + kBlock - 1) / kBlock + 1;
  const unsigned long long cap = (unsigned long long)c.sm_count * 16;
  const unsigned grid = (unsigned)(want < cap ? want : cap);
  if (big_endian) k_well_formed_utf16<true><<<grid, kBlock, 0, c.stream>>>(in, len, out);
  else k_well_formed_utf16<false><<<grid, kBlock, 0, c.stream>>>(in, len, out);
  count_launch(1);
  return cudaGetLastError();
}

cudaError_t launch_change_endianness_utf16(const LaunchCtx &c, const uint16_t *in, size_t len, uint16_t *out) {
  const unsigned long long want = (len / 8 + kBlock - 1) / kBlock + 1;
  const unsigned long long cap = (unsigned long long)c.sm_count * 16;
  k_swap_utf16<<<(unsigned)(want < cap ? want : cap), kBlock, 0, c.stream>>>(in, len, out);
  count_launch(1);
  return cudaGetLastError();
}

}  // namespace b200
