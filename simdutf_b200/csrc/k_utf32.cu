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
#include <cstdlib>
#include <type_traits>

#include "bp_device.cuh"
#include "device_common.cuh"
#include "launch.h"

namespace b200 {

namespace {

using bpd::kChunkTiles;
using bpd::kThreads;
using bpd::kWarpsPerCta;

constexpr uint32_t kTileBytes = 2048u;  // 64 input bytes per lane

__device__ __forceinline__ InView make_view_elems(const void *p, size_t bytes) {
  InView v;
  const uintptr_t a = reinterpret_cast<uintptr_t>(p);
  v.base = reinterpret_cast<const uint4 *>(a & ~uintptr_t(15));
  v.vbeg = a & 15u;
  v.vend = v.vbeg + bytes;
  return v;
}
__device__ __forceinline__ uint32_t bswap16(uint32_t u) { return ((u >> 8) | (u << 8)) & 0xFFFFu; }

// ---- per-element rules -----------------------------------------------------------------------------------
// count(v, nv): output elements of input element v; emit(): the same count, the elements packed little-endian
// into P (first element lowest; OutT-sized fields), and the element's own error code.
struct U32ToU8 {
  using In = uint32_t;
  using Out = uint8_t;
  static constexpr uint32_t kMax = 4;
  static constexpr bool kNeedsNeighbours = false;
  __device__ static uint32_t count(uint32_t w, uint32_t, uint32_t) { return 1u + (w > 0x7Fu) + (w > 0x7FFu) + (w > 0xFFFFu); }
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
  __device__ static uint32_t count(uint32_t w, uint32_t, uint32_t) { return 1u + (w > 0xFFFFu); }
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
  // one code point per unit that is not a low surrogate (== count_utf16 / utf32_length_from_utf16, so a buffer
  // sized by that query is never overrun); a high surrogate emits the pair's code point, looking one unit ahead
  __device__ static uint32_t count(uint32_t u, uint32_t, uint32_t) {
    if (BE) u = bswap16(u);
    return (u & 0xFC00u) != 0xDC00u;
  }
  __device__ static uint32_t emit(uint32_t u, uint32_t pu, uint32_t nu, bool has_prev, bool has_next, uint32_t &P, int &err) {
    if (BE) { u = bswap16(u); pu = bswap16(pu); nu = bswap16(nu); }
    err = u16_bad(u, pu, has_prev, nu, has_next) ? kSurrogate : kSuccess;
    if ((u & 0xFC00u) == 0xDC00u) return 0u;
    P = (u & 0xFC00u) == 0xD800u ? 0x10000u + ((u - 0xD800u) << 10) + ((nu - 0xDC00u) & 0x3FFu) : u;
    return 1u;
  }
};

template <class T>
struct Shape {
  using In = typename T::In;
  using Out = typename T::Out;
  static constexpr uint32_t kInPerLane = 64u / sizeof(In);            // 16 code points or 32 units
  static constexpr uint32_t kVec = 16u / sizeof(Out);                 // output elements per 16-byte vector
  static constexpr uint32_t kMaxOut = kInPerLane * T::kMax;           // per lane
  static constexpr uint32_t kStrideWords = (((kMaxOut + kVec) * sizeof(Out) + 3u) / 4u) | 1u;
  static constexpr uint32_t kSmemBytes = kWarpsPerCta * 32u * kStrideWords * 4u;
  static constexpr uint32_t kMaxVec = (kMaxOut + kVec - 1u) / kVec;
};

// Element i (virtual index from the aligned base) of the input; zero outside the buffer.
template <class In>
__device__ __forceinline__ uint32_t elem_guarded(const InView &in, long long i) {
  const long long pos = i * (long long)sizeof(In);
  if (pos < (long long)in.vbeg || pos >= (long long)in.vend) return 0u;
  return (uint32_t)__ldg(reinterpret_cast<const In *>(in.base) + i);
}

// This lane's 64 input bytes as elements (zero filler outside the buffer).
template <class In>
__device__ __forceinline__ void load_lane(const InView &in, unsigned long long r0, bool interior, uint32_t (&w)[16]) {
  if (interior) {
    const uint4 *gp = in.base + (r0 >> 4);
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const uint4 v = __ldg(gp + j);
      w[4 * j] = v.x; w[4 * j + 1] = v.y; w[4 * j + 2] = v.z; w[4 * j + 3] = v.w;
    }
  } else {
#pragma unroll
    for (int j = 0; j < 4; j++) {
      bool ins;
      load_granule(in, (r0 >> 4) + (unsigned long long)j, &w[4 * j], ins);
    }
  }
}
template <class In>
__device__ __forceinline__ uint32_t lane_elem(const uint32_t (&w)[16], int i) {
  return sizeof(In) == 4 ? w[i] : (w[i >> 1] >> (16 * (i & 1))) & 0xFFFFu;
}

// ---- counts pass -----------------------------------------------------------------------------------------
template <class T>
__global__ void __launch_bounds__(kThreads) k_elem_tile_counts(const void *ptr, size_t bytes, uint16_t *tile_cnt,
                                                                unsigned long long *chunk_off, uint32_t num_tiles,
                                                                uint32_t num_chunks, Scratch *scr) {
  using S = Shape<T>;
  using In = typename T::In;
  const InView in = make_view_elems(ptr, bytes);
  const unsigned lane = threadIdx.x & 31u;
  bpd::counts_pass(
      [&](uint32_t t) -> uint32_t {
        const unsigned long long t0 = (unsigned long long)t * kTileBytes, r0 = t0 + lane * 64ull;
        const bool interior = t0 >= in.vbeg && t0 + kTileBytes <= in.vend;
        uint32_t w[16];
        load_lane<In>(in, r0, interior, w);
        uint32_t cnt = 0;
#pragma unroll
        for (int i = 0; i < (int)S::kInPerLane; i++) {
          const unsigned long long pos = r0 + (unsigned long long)i * sizeof(In);
          const uint32_t c = T::count(lane_elem<In>(w, i), 0u, 0u);
          cnt += (interior || (pos >= in.vbeg && pos < in.vend)) ? c : 0u;
        }
        return bpd::warp_sum_u32(cnt);
      },
      tile_cnt, chunk_off, num_tiles, num_chunks, scr);
}

// ---- emit pass -------------------------------------------------------------------------------------------
template <class Out>
__device__ __forceinline__ void sts_elem(uint32_t addr, uint32_t v) {
  if (sizeof(Out) == 1) bpd::sts_u8(addr, v);
  else if (sizeof(Out) == 2) bpd::sts_u16(addr, v);
  else bpd::sts_u32(addr, v);
}

template <class T, int MINB>
__global__ void __launch_bounds__(kThreads, MINB)
k_elem_transcode(const void *ptr, size_t bytes, typename T::Out *out, const uint16_t *tile_cnt,
                 const unsigned long long *chunk_off, uint32_t num_tiles, uint32_t num_chunks, Scratch *scr,
                 ResultPOD *res) {
  using S = Shape<T>;
  using In = typename T::In;
  using Out = typename T::Out;
  extern __shared__ __align__(16) uint32_t smem[];
  const InView in = make_view_elems(ptr, bytes);
  const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  const uint32_t nwarps = gridDim.x * kWarpsPerCta;
  uint32_t *region_w = smem + (warp * 32u + lane) * S::kStrideWords;
  Out *region = reinterpret_cast<Out *>(region_w);
  const unsigned long long out_elems = (unsigned long long)(reinterpret_cast<uintptr_t>(out) / sizeof(Out));
  const long long first_elem = (long long)(in.vbeg / sizeof(In)), end_elem = (long long)(in.vend / sizeof(In));

  for (uint32_t tile = blockIdx.x * kWarpsPerCta + warp; tile < num_tiles; tile += nwarps) {
    const unsigned long long t0 = (unsigned long long)tile * kTileBytes, r0 = t0 + lane * 64ull;
    const bool interior = t0 >= in.vbeg + 16ull && t0 + kTileBytes + 16ull <= in.vend;
    const uint32_t before = bpd::tile_before_partial(tile_cnt, tile);
    const unsigned long long coff = chunk_off[tile / kChunkTiles];
    uint32_t w[16];
    load_lane<In>(in, r0, interior, w);
    const long long e0 = (long long)(r0 / sizeof(In));  // virtual index of this lane's first element
    uint32_t pv = 0, nv = 0;
    if (T::kNeedsNeighbours) {
      pv = elem_guarded<In>(in, e0 - 1);
      nv = elem_guarded<In>(in, e0 + (long long)S::kInPerLane);
    }
    const unsigned long long goff = coff + bpd::warp_sum_u32(before);

    // per-element outputs, kept packed; count first (the lane's offset decides where its region starts)
    uint32_t P[S::kInPerLane];
    uint32_t n[S::kInPerLane];
    uint32_t cnt = 0;
    long long bad_at = -1;
    int bad_code = 0;
#pragma unroll
    for (int i = 0; i < (int)S::kInPerLane; i++) {
      const long long idx = e0 + i;
      const bool inside = interior || (idx >= first_elem && idx < end_elem);
      const uint32_t v = lane_elem<In>(w, i);
      const uint32_t p = i ? lane_elem<In>(w, i - 1) : pv;
      const uint32_t nx = i + 1 < (int)S::kInPerLane ? lane_elem<In>(w, i + 1 < (int)S::kInPerLane ? i + 1 : i) : nv;
      int err;
      uint32_t c = T::emit(v, p, nx, idx > first_elem, idx + 1 < end_elem, P[i], err);
      if (!inside) { c = 0; err = 0; }
      if (err && bad_at < 0) { bad_at = idx; bad_code = err; }
      n[i] = c;
      cnt += c;
    }
    if (bad_at >= 0) {
      const unsigned long long key = err_key((unsigned long long)(bad_at - first_elem), bad_code);
      if (key < ld_relaxed_u64(&scr->err_key)) report_error(scr, key);
    }
    const uint32_t incl = bpd::warp_inclusive_u32(cnt);
    const unsigned long long G = goff + (incl - cnt);
    const uint32_t a = (uint32_t)((out_elems + G) & (S::kVec - 1u));

    // compaction into the private region
    {
      uint32_t sp = (uint32_t)__cvta_generic_to_shared(region + a);
#pragma unroll
      for (int i = 0; i < (int)S::kInPerLane; i++) {
#pragma unroll
        for (uint32_t k = 0; k < T::kMax; k++) {
          if (k < n[i]) sts_elem<Out>(sp + k * (uint32_t)sizeof(Out), sizeof(Out) == 4 ? P[i] : P[i] >> (8u * (uint32_t)sizeof(Out) * k));
        }
        sp += n[i] * (uint32_t)sizeof(Out);
      }
    }
    __syncwarp();

    // staging -> global: a lane owns the 16-byte vectors that hold its elements except its last partial one, which
    // the lane to its right completes (it copies the elements in front of its own first one from this lane's tail)
    {
      Out *gbase = out + G - a;
      const uint32_t end = a + cnt;
      if (__all_sync(kFull, cnt >= S::kVec)) {
        const uint32_t prev_end = __shfl_up_sync(kFull, end, 1);
        if (lane > 0) {
          const Out *src = region - S::kStrideWords * (4u / (uint32_t)sizeof(Out)) + (prev_end - a);
#pragma unroll
          for (uint32_t u = 0; u + 1 < S::kVec; u++)
            if (u < a) region[u] = src[u];
        }
        const uint32_t vfull = end / S::kVec;
        uint32_t v0 = 0;
        if (lane == 0 && a > 0) {
#pragma unroll
          for (uint32_t u = 1; u < S::kVec; u++)
            if (u >= a) gbase[u] = region[u];
          v0 = 1;
        }
        if (lane == 31) {
#pragma unroll
          for (uint32_t u = 0; u + 1 < S::kVec; u++) {
            const uint32_t i = vfull * S::kVec + u;
            if (i < end) gbase[i] = region[i];
          }
        }
#pragma unroll
        for (uint32_t v = 0; v < S::kMaxVec; v++) {
          if (v >= v0 && v < vfull) {
            uint4 x;
            x.x = region_w[4u * v];
            x.y = region_w[4u * v + 1u];
            x.z = region_w[4u * v + 2u];
            x.w = region_w[4u * v + 3u];
            stg_stream_v4(reinterpret_cast<uint4 *>(gbase) + v, x);
          }
        }
      } else {
        for (uint32_t i = a; i < end; i++) gbase[i] = region[i];
      }
    }
    __syncwarp();
  }

  if (grid_last_thread(scr)) {
    bpd::write_result_from_key(res, ld_relaxed_u64(&scr->err_key), chunk_off[num_chunks]);
    scratch_reset(scr);
  }
}

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
  for (size_t v = tid; v < nvec; v += nthreads) {
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

inline size_t tiles_for(const void *in, size_t bytes) {
  const size_t span = (reinterpret_cast<uintptr_t>(in) & 15u) + bytes;
  return (span + kTileBytes - 1) / kTileBytes;
}
inline size_t workspace_slots(size_t tiles) {
  const size_t chunks = (tiles + kChunkTiles - 1) / kChunkTiles;
  return chunks + 1 + (tiles * sizeof(uint16_t) + 7) / 8 + 1;
}

template <class T>
cudaError_t launch_elem(const LaunchCtx &c, const void *in, size_t len, void *out, void *res) {
  using S = Shape<T>;
  constexpr int MINB = 2;
  const size_t bytes = len * sizeof(typename T::In);
  const size_t tiles = tiles_for(in, bytes);
  if (workspace_slots(tiles) > c.desc_capacity || tiles > 0xFFFFFF00ull) return cudaErrorInvalidValue;
  static int per_sm = 0;
  if (per_sm == 0) {
    cudaError_t e = cudaFuncSetAttribute(k_elem_transcode<T, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)S::kSmemBytes);
    if (e != cudaSuccess) return e;
    int n = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k_elem_transcode<T, MINB>, kThreads, S::kSmemBytes);
    if (e != cudaSuccess) return e;
    per_sm = n < 1 ? 1 : n;
  }
  const size_t chunks = (tiles + kChunkTiles - 1) / kChunkTiles;
  unsigned long long *chunk_off = c.desc;
  uint16_t *tile_cnt = reinterpret_cast<uint16_t *>(c.cnt);
  {
    const size_t cap = (size_t)c.sm_count * 8;
    const unsigned grid = (unsigned)(chunks < cap ? chunks : cap);
    k_elem_tile_counts<T><<<grid, kThreads, 0, c.stream>>>(in, bytes, tile_cnt, chunk_off, (uint32_t)tiles, (uint32_t)chunks, c.scratch);
  }
  {
    const size_t ctas = (tiles + kWarpsPerCta - 1) / kWarpsPerCta;
    const size_t cap = (size_t)c.sm_count * per_sm;
    const unsigned grid = (unsigned)(ctas < cap ? ctas : cap);
    k_elem_transcode<T, MINB><<<grid, kThreads, S::kSmemBytes, c.stream>>>(
        in, bytes, static_cast<typename T::Out *>(out), tile_cnt, chunk_off, (uint32_t)tiles, (uint32_t)chunks, c.scratch,
        static_cast<ResultPOD *>(res));
  }
  count_launch(2);
  return cudaGetLastError();
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

}  // namespace b200
