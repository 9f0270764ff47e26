// k_utf32.cu — the UTF-32 side of SURVEY.md §8f rank 1 on sm_100a:
//   validate_utf32[_with_errors]                          (reference src/scalar/utf32.h:10-38)
//   utf8_length_from_utf32 / utf16_length_from_utf32       (src/scalar/utf32.h:40-66)
//   convert_utf32_to_utf8[_with_errors]                   (src/scalar/utf32_to_utf8/utf32_to_utf8.h:63-124)
//   convert_utf32_to_utf16le/be[_with_errors]             (src/scalar/utf32_to_utf16/utf32_to_utf16.h:52-86)
//   convert_utf16le/be_to_utf32[_with_errors]             (src/scalar/utf16_to_utf32/utf16_to_utf32.h:42-76)
//
// With 16- and 32-bit input elements the per-element form is already cheap per input BYTE (4-12 logic ops per element),
// so these kernels do not transpose to bit planes; they share the transcoders' data path instead: a counts pass
// (per-tile output counts -> chunk offsets, bp_device.cuh) and a warp-independent emit pass in which every lane owns
// 64 contiguous input bytes, compacts its output into a lane-private staging region (odd word stride, aligned to the
// destination's 16-byte vectors) and streams its own vectors out.
// Every element is judged on its own (UTF-16 input: together with its two neighbours), so the first error is the
// atomicMin of (index << 8 | code).
#include "elem_device.cuh"

namespace b200 {

namespace {

using namespace elem;

// ---- per-element rules -----------------------------------------------------------------------------------
// emit(): the number of output elements of input element v (what the conversion's length query counts), the elements
// packed little-endian into P (first element lowest; OutT-sized fields), and the element's own error code.
struct U32ToU8 {
  using In = uint32_t;
  using Out = uint8_t;
  static constexpr uint32_t kMax = 4;
  static constexpr bool kNeedsNeighbours = false;
  static constexpr bool kFast = false;
  // (a select-only form of this was measured: 13 % slower — the four candidate encodings cost more than the divergence)
  __device__ static uint32_t emit(uint32_t w, uint32_t, uint32_t, bool, bool, uint32_t &P, int &err) {
    err = w > 0x10FFFFu ? kTooLarge : ((w & 0xFFFFF800u) == 0xD800u ? kSurrogate : kSuccess);
    const uint32_t c0 = w & 0x3Fu, c1 = (w >> 6) & 0x3Fu, c2 = (w >> 12) & 0x3Fu;
    if (w <= 0x7Fu) { P = w; return 1u; }
    if (w <= 0x7FFu) { P = (0xC0u | (w >> 6)) | ((0x80u | c0) << 8); return 2u; }
    if (w <= 0xFFFFu) { P = (0xE0u | (w >> 12)) | ((0x80u | c1) << 8) | ((0x80u | c0) << 16); return 3u; }
    P = (0xF0u | ((w >> 18) & 7u)) | ((0x80u | c2) << 8) | ((0x80u | c1) << 16) | ((0x80u | c0) << 24);
    return 4u;
  }
};
template <bool BE>
struct U32ToU16 {
  using In = uint32_t;
  using Out = uint16_t;
  static constexpr uint32_t kMax = 2;
  static constexpr bool kNeedsNeighbours = false;
  static constexpr bool kFast = false;
  __device__ static uint32_t emit(uint32_t w, uint32_t, uint32_t, bool, bool, uint32_t &P, int &err) {
    err = w > 0x10FFFFu ? kTooLarge : ((w & 0xFFFFF800u) == 0xD800u ? kSurrogate : kSuccess);
    if (w <= 0xFFFFu) {
      P = BE ? bswap16(w) : w;
      return 1u;
    }
    const uint32_t c = w - 0x10000u;
    const uint32_t hi = 0xD800u | ((c >> 10) & 0x3FFu), lo = 0xDC00u | (c & 0x3FFu);
    P = BE ? (bswap16(hi) | (bswap16(lo) << 16)) : (hi | (lo << 16));
    return 2u;
  }
};
template <bool BE>
struct U16ToU32 {
  using In = uint16_t;
  using Out = uint32_t;
  static constexpr uint32_t kMax = 1;
  static constexpr bool kNeedsNeighbours = true;
  static constexpr bool kFast = false;
  // one code point per unit that is not a low surrogate (== count_utf16 / utf32_length_from_utf16, so a buffer
  // sized by that query is never overrun); a high surrogate emits the pair's code point, looking one unit ahead
  __device__ static uint32_t emit(uint32_t u, uint32_t pu, uint32_t nu, bool has_prev, bool has_next, uint32_t &P, int &err) {
    if (BE) { u = bswap16(u); pu = bswap16(pu); nu = bswap16(nu); }
    const bool low = (u & 0xFC00u) == 0xDC00u, high = (u & 0xFC00u) == 0xD800u;
    const bool bad = (low && !(has_prev && (pu & 0xFC00u) == 0xD800u)) || (high && !(has_next && (nu & 0xFC00u) == 0xDC00u));
    err = bad ? kSurrogate : kSuccess;
    const uint32_t pair = 0x10000u + ((u - 0xD800u) << 10) + ((nu - 0xDC00u) & 0x3FFu);
    P = high ? pair : u;
    return low ? 0u : 1u;
  }
};

// ---- reductions: validate_utf32, utf8/utf16 length from utf32 ---------------------------------------------
// MODE 0: validate (first error), 1: utf8_length_from_utf32, 2: utf16_length_from_utf32
template <int MODE>
__global__ void __launch_bounds__(kBlock) k_scan_utf32(const uint32_t *in, size_t len, Scratch *scr, void *out) {
  __shared__ unsigned long long s_part[kWarps];
  const size_t tid = (size_t)blockIdx.x * kBlock + threadIdx.x, nthreads = (size_t)gridDim.x * kBlock;
  const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  unsigned long long total = 0;
  unsigned long long best = kNoError;
  // 16-byte vectors where the pointer allows it, single words at the ragged ends
  const uintptr_t a = reinterpret_cast<uintptr_t>(in);
  size_t head = ((16u - (a & 15u)) & 15u) >> 2;
  if (head > len) head = len;
  const size_t nvec = (len - head) >> 2;
  const uint4 *vi = reinterpret_cast<const uint4 *>(in + head);
  auto one = [&](uint32_t w, size_t idx) {
    if (MODE == 0) {
      const int code = w > 0x10FFFFu ? kTooLarge : ((w & 0xFFFFF800u) == 0xD800u ? kSurrogate : kSuccess);
      if (code) {
        const unsigned long long k = err_key(idx, code);
        best = k < best ? k : best;
      }
    } else if (MODE == 1) {
      total += 1u + (w > 0x7Fu) + (w > 0x7FFu) + (w > 0xFFFFu);
    } else {
      total += 1u + (w > 0xFFFFu);
    }
  };
  // four vectors per thread and round, all four loads in flight before the first is used (one load per round left the
  // kernel latency-bound: 0.68 of the copy bandwidth).  Validation folds the 16 words of a round into two extremes —
  // the largest word and the smallest (word ^ 0xD800), which is < 0x800 exactly for a surrogate (the reference keeps a
  // running maximum the same way, src/icelake/icelake_utf32_validation.inl.cpp) — and looks at single words only in a
  // round that tripped one of them.
  size_t v = tid;
  bool stop = false;  // warp-uniform
  for (; (v - lane) + 31u + 3 * nthreads < nvec && !stop; v += 4 * nthreads) {  // whole warps only: the round votes
    uint4 x[4];
#pragma unroll
    for (int k = 0; k < 4; k++) x[k] = ldg_stream_v4(vi + v + (size_t)k * nthreads);
    if (MODE == 0) {
      uint32_t mx = 0u, mn = 0xFFFFFFFFu;
#pragma unroll
      for (int k = 0; k < 4; k++) {
        mx = max(max(mx, x[k].x), max(x[k].y, max(x[k].z, x[k].w)));
        mn = min(min(mn, x[k].x ^ 0xD800u), min(x[k].y ^ 0xD800u, min(x[k].z ^ 0xD800u, x[k].w ^ 0xD800u)));
      }
      if (__any_sync(kFull, mx > 0x10FFFFu || mn < 0x800u)) {
        // An error in front of this round that is already on record ends the warp's work: the first error wins and the
        // warp's indices only grow.  Otherwise the round's first error goes on record NOW (one atomic per warp), so that
        // text in another encoding (detect_encodings) is given up after one round instead of being searched word by word.
        unsigned long long cur = ld_relaxed_u64(&scr->err_key);
        cur = __shfl_sync(kFull, cur, 0);
        if (cur != kNoError && (cur >> 8) < head + 4 * (v - lane)) {
          stop = true;
        } else {
#pragma unroll
          for (int k = 0; k < 4; k++) {
            const size_t i = head + 4 * (v + (size_t)k * nthreads);
            one(x[k].x, i); one(x[k].y, i + 1); one(x[k].z, i + 2); one(x[k].w, i + 3);
          }
          const unsigned long long wbest = warp_min_u64(best);
          if (lane == 0 && wbest != kNoError) report_error(scr, wbest);
        }
      }
    } else {
#pragma unroll
      for (int k = 0; k < 4; k++) {
        const size_t i = head + 4 * (v + (size_t)k * nthreads);
        one(x[k].x, i); one(x[k].y, i + 1); one(x[k].z, i + 2); one(x[k].w, i + 3);
      }
    }
  }
  for (; v < nvec && !stop; v += nthreads) {
    const uint4 x = ldg_stream_v4(vi + v);
    const size_t i = head + 4 * v;
    one(x.x, i); one(x.y, i + 1); one(x.z, i + 2); one(x.w, i + 3);
  }
  for (size_t i = tid; i < head; i += nthreads) one(in[i], i);
  for (size_t i = head + 4 * nvec + tid; i < len; i += nthreads) one(in[i], i);
  if (MODE == 0) {
    best = warp_min_u64(best);
    if (lane == 0 && best != kNoError) report_error(scr, best);
  } else {
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
    if (MODE == 0) bpd::write_result_from_key(static_cast<ResultPOD *>(out), ld_relaxed_u64(&scr->err_key), len);
    else *static_cast<unsigned long long *>(out) = ld_relaxed_u64(&scr->acc0);
    scratch_reset(scr);
  }
}

// detect_encodings (reference src/fallback/implementation.cpp:8-32, src/encoding_types.cpp:32-49): a BOM is trusted;
// otherwise the answer is the set of encodings the buffer validates as.  The three validators have already run on
// this stream into r8 / r16 / r32 (the latter two only when the length allows); this combines them.
__global__ void k_detect_finish(const uint8_t *in, size_t len, const ResultPOD *r8, const ResultPOD *r16, const ResultPOD *r32,
                                unsigned long long *out) {
  const uint32_t b0 = len > 0 ? in[0] : 0x100u, b1 = len > 1 ? in[1] : 0x100u, b2 = len > 2 ? in[2] : 0x100u,
                 b3 = len > 3 ? in[3] : 0x100u;
  unsigned long long ans = 0;
  if (b0 == 0xFFu && b1 == 0xFEu) ans = (b2 == 0u && b3 == 0u) ? 8u : 2u;                         // UTF32_LE : UTF16_LE
  else if (b0 == 0xFEu && b1 == 0xFFu) ans = 4u;                                                   // UTF16_BE
  else if (b0 == 0u && b1 == 0u && b2 == 0xFEu && b3 == 0xFFu) ans = 16u;                          // UTF32_BE
  else if (len >= 4 && b0 == 0xEFu && b1 == 0xBBu && b2 == 0xBFu) ans = 1u;                        // UTF8
  else {
    if (r8->error == kSuccess) ans |= 1u;
    if ((len & 1u) == 0 && r16->error == kSuccess) ans |= 2u;
    if ((len & 3u) == 0 && r32->error == kSuccess) ans |= 8u;
  }
  *out = ans;
}

}  // namespace

size_t utf32_family_tiles(const void *in, size_t bytes) { return workspace_slots(tiles_for(in, bytes)); }

cudaError_t launch_convert_utf32_to_utf8(const LaunchCtx &c, const uint32_t *in, size_t len, char *out, void *res) {
  return launch_elem<U32ToU8>(c, in, len, out, res);
}
cudaError_t launch_convert_utf32_to_utf16(const LaunchCtx &c, const uint32_t *in, size_t len, uint16_t *out, void *res,
                                          bool big_endian) {
  return big_endian ? launch_elem<U32ToU16<true>>(c, in, len, out, res) : launch_elem<U32ToU16<false>>(c, in, len, out, res);
}
cudaError_t launch_convert_utf16_to_utf32(const LaunchCtx &c, const uint16_t *in, size_t len, uint32_t *out, void *res,
                                          bool big_endian) {
  return big_endian ? launch_elem<U16ToU32<true>>(c, in, len, out, res) : launch_elem<U16ToU32<false>>(c, in, len, out, res);
}
cudaError_t launch_scan_utf32(const LaunchCtx &c, const uint32_t *in, size_t len, void *out, int mode) {
  const unsigned long long want = (len / 4 + kBlock - 1) / kBlock + 1;
  const unsigned long long cap = (unsigned long long)c.sm_count * 8;
  const unsigned grid = (unsigned)(want < cap ? want : cap);
  if (mode == 0) k_scan_utf32<0><<<grid, kBlock, 0, c.stream>>>(in, len, c.scratch, out);
  else if (mode == 1) k_scan_utf32<1><<<grid, kBlock, 0, c.stream>>>(in, len, c.scratch, out);
  else k_scan_utf32<2><<<grid, kBlock, 0, c.stream>>>(in, len, c.scratch, out);
  count_launch(1);
  return cudaGetLastError();
}

cudaError_t launch_detect_finish(const LaunchCtx &c, const char *in, size_t len, const void *r8, const void *r16, const void *r32,
                                 unsigned long long *out) {
  k_detect_finish<<<1, 1, 0, c.stream>>>(reinterpret_cast<const uint8_t *>(in), len, static_cast<const ResultPOD *>(r8),
                                         static_cast<const ResultPOD *>(r16), static_cast<const ResultPOD *>(r32), out);
  count_launch(1);
  return cudaGetLastError();
}

}  // namespace b200
