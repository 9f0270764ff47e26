// k_latin1.cu — the Latin-1 / ASCII side of SURVEY.md §8f rank 3 on sm_100a:
//   validate_ascii[_with_errors]                (reference src/scalar/ascii.h:36-64)
//   utf8_length_from_latin1                     (src/scalar/latin1.h:9-19; src/fallback/implementation.cpp:431-465)
//   convert_latin1_to_utf8                      (src/scalar/latin1_to_utf8/latin1_to_utf8.h:9-46)
//   convert_latin1_to_utf16le/be, _to_utf32     (src/scalar/latin1_to_utf16/latin1_to_utf16.h, latin1_to_utf32.h)
//   convert_utf8_to_latin1[_with_errors]        (src/scalar/utf8_to_latin1/utf8_to_latin1.h:83-149)
//   convert_utf16le/be_to_latin1[_with_errors]  (src/scalar/utf16_to_latin1/utf16_to_latin1.h:38-92)
//   convert_utf32_to_latin1[_with_errors]       (src/scalar/utf32_to_latin1/utf32_to_latin1.h:33-62)
//
// Two shapes.  The conversions whose output index differs from the input index (Latin-1 <-> UTF-8) are traits of
// the per-element transcoder (elem_device.cuh: counts pass + warp-independent emit pass).  The others are maps
// (output element i depends on input element i only): one kernel, one 16-byte store per thread and step.
// UTF-8 -> Latin-1 is judged per byte with its two neighbours; this is exactly the reference's sequential rule
// because a valid Latin-1-range text has only 1- and 2-byte characters, so whether a continuation byte is
// "consumed" depends on the byte before it alone (see U8ToL1::emit).
#include "elem_device.cuh"

namespace b200 {

namespace {

using namespace elem;

struct L1ToU8 {
  using In = uint8_t;
  using Out = uint8_t;
  static constexpr uint32_t kMax = 2;
  static constexpr bool kNeedsNeighbours = false;
  static constexpr bool kFast = true;
  // 64 + the number of bytes >= 0x80: the high bits summed as packed byte counters (<= 16 each)
  __device__ static uint32_t fast_pass1(const uint32_t (&w)[16], uint32_t, uint32_t, bool &bad) {
    bad = false;
    return 64u + l1_high_count64(w);  // swar.h (host-tested)
  }
  // four input bytes -> 4..8 staged bytes (tiles inside the buffer)
  __device__ static void emit_word(uint32_t w, uint32_t, uint32_t &sp) {
    uint32_t f, c;
    const uint32_t hi = l1u8_word(w, &f, &c);
#pragma unroll
    for (int k = 0; k < 4; k++) {
      const uint32_t h = (hi >> (8 * k)) & 1u;
      bpd::sts_u8(sp, f >> (8 * k));
      if (h) bpd::sts_u8(sp + 1u, c >> (8 * k));
      sp += 1u + h;
    }
  }
  __device__ static uint32_t emit(uint32_t b, uint32_t, uint32_t, bool, bool, uint32_t &P, int &err) {
    err = kSuccess;
    const uint32_t hi = b >> 7;
    P = hi ? ((0xC0u | (b >> 6)) | ((0x80u | (b & 0x3Fu)) << 8)) : b;
    return 1u + hi;
  }
};

// One Latin-1 byte per non-continuation byte (== count_utf8 == latin1_length_from_utf8, so a buffer sized by that
// query is never overrun).  The reference walks characters; with only 1- and 2-byte characters allowed the walk
// reaches byte p as a character start unless p is a continuation byte directly behind a C0..DF lead, so:
//   80..BF  not behind a C0..DF lead -> TOO_LONG         C0..DF  no continuation behind it -> TOO_SHORT,
//   E0..F7  -> TOO_LARGE, F8..FF -> HEADER_BITS                  else C0/C1 -> OVERLONG, C4..DF -> TOO_LARGE
// and the first error of the walk is the smallest flagged position (everything before it is well formed).
struct U8ToL1 {
  using In = uint8_t;
  using Out = uint8_t;
  static constexpr uint32_t kMax = 1;
  static constexpr bool kNeedsNeighbours = true;
  static constexpr bool kFast = true;
  // 64 - the number of continuation bytes; `bad` (conservative) unless every byte >= 0xC0 is C2 / C3 and the
  // continuation bytes are exactly the bytes behind those leads (the lane's neighbours pb / nb included)
  __device__ static uint32_t fast_pass1(const uint32_t (&w)[16], uint32_t pb, uint32_t nb, bool &bad) {
    return 64u - u8l1_screen64(w, pb, nb, &bad);  // swar.h (host-tested)
  }
  // four input bytes of a screened lane -> 0..4 staged bytes; wn: the word behind them (its low byte is all that is used)
  __device__ static void emit_word(uint32_t w, uint32_t wn, uint32_t &sp) {
    uint32_t keep7;
    const uint32_t o = u8l1_word(w, wn, &keep7);
    const uint32_t k0 = (keep7 >> 7) & 1u, k1 = (keep7 >> 15) & 1u, k2 = (keep7 >> 23) & 1u, k3 = keep7 >> 31;
    if (k0) bpd::sts_u8(sp, o);
    if (k1) bpd::sts_u8(sp + k0, o >> 8);
    if (k2) bpd::sts_u8(sp + k0 + k1, o >> 16);
    if (k3) bpd::sts_u8(sp + k0 + k1 + k2, o >> 24);
    sp += k0 + k1 + k2 + k3;
  }
  // branch-free; when the caller ignores err (tiles screened by fast_pass1) only the two selects for P and the count remain
  __device__ static uint32_t emit(uint32_t b, uint32_t pb, uint32_t nb, bool has_prev, bool has_next, uint32_t &P, int &err) {
    const bool cont = (b & 0xC0u) == 0x80u, lead2 = (b & 0xE0u) == 0xC0u, big = b >= 0xE0u;
    const bool next_cont = has_next && (nb & 0xC0u) == 0x80u, prev_lead = has_prev && (pb & 0xE0u) == 0xC0u;
    int e = kSuccess;
    e = (cont && !prev_lead) ? kTooLong : e;
    e = lead2 ? (!next_cont ? kTooShort : (b < 0xC2u ? kOverlong : (b > 0xC3u ? kTooLarge : kSuccess))) : e;
    e = big ? (b < 0xF8u ? kTooLarge : kHeaderBits) : e;
    err = e;
    P = lead2 ? (((b & 3u) << 6) | (nb & 0x3Fu)) : b;
    return cont ? 0u : 1u;
  }
};

// ---- maps: Latin-1 -> UTF-16LE/BE / UTF-32, UTF-16LE/BE / UTF-32 -> Latin-1 ------------------------------------
// A thread produces one aligned 16-byte vector of output per step from 16 / sizeof(Out) input elements (vector
// loads when the input pointer allows, element loads otherwise); the ragged head and tail are element-wise.
template <class In, class Out, bool SWAP_IN, bool SWAP_OUT>
__device__ __forceinline__ uint32_t map_one(uint32_t v, bool &bad) {
  if (SWAP_IN) v = bswap16(v);
  if (sizeof(Out) == 1) bad = bad || v > 0xFFu;  // narrowing: TOO_LARGE
  if (SWAP_OUT) v = bswap16(v);
  return v;
}

template <class In, class Out, bool SWAP_IN, bool SWAP_OUT>
__global__ void __launch_bounds__(kBlock) k_latin1_map(const In *in, size_t len, Out *out, Scratch *scr, ResultPOD *res) {
  constexpr uint32_t kG = 16u / (uint32_t)sizeof(Out);       // elements per output vector
  constexpr uint32_t kInBytes = kG * (uint32_t)sizeof(In);   // 4, 8, 32 or 64
  constexpr uint32_t kInAlign = kInBytes > 16u ? 16u : kInBytes;
  const size_t tid = (size_t)blockIdx.x * kBlock + threadIdx.x, nthreads = (size_t)gridDim.x * kBlock;
  size_t head = ((16u - (reinterpret_cast<uintptr_t>(out) & 15u)) & 15u) / sizeof(Out);
  if (head > len) head = len;
  const size_t ngroups = (len - head) / kG;
  const In *gin = in + head;
  uint4 *gout = reinterpret_cast<uint4 *>(out + head);
  const bool in_vec = (reinterpret_cast<uintptr_t>(gin) & (kInAlign - 1u)) == 0;
  unsigned long long best = kNoError;

  auto element = [&](size_t i) {
    bool bad = false;
    const uint32_t v = map_one<In, Out, SWAP_IN, SWAP_OUT>((uint32_t)in[i], bad);
    out[i] = (Out)v;
    if (bad) {
      const unsigned long long k = err_key(i, kTooLarge);
      best = k < best ? k : best;
    }
  };

  for (size_t g = tid; g < ngroups; g += nthreads) {
    uint32_t e[kG];
    const In *p = gin + g * kG;
    if (in_vec) {
      uint32_t w[kInBytes / 4u];
      if constexpr (kInBytes == 4u) {
        w[0] = __ldg(reinterpret_cast<const uint32_t *>(p));
      } else if constexpr (kInBytes == 8u) {
        const uint2 x = __ldg(reinterpret_cast<const uint2 *>(p));
        w[0] = x.x; w[1] = x.y;
      } else {
#pragma unroll
        for (uint32_t j = 0; j < kInBytes / 16u; j++) {
          const uint4 x = ldg_stream_v4(reinterpret_cast<const uint4 *>(p) + j);
          w[4 * j] = x.x; w[4 * j + 1] = x.y; w[4 * j + 2] = x.z; w[4 * j + 3] = x.w;
        }
      }
#pragma unroll
      for (uint32_t k = 0; k < kG; k++) {
        if (sizeof(In) == 4) e[k] = w[k];
        else if (sizeof(In) == 2) e[k] = (w[k >> 1] >> (16u * (k & 1u))) & 0xFFFFu;
        else e[k] = (w[k >> 2] >> (8u * (k & 3u))) & 0xFFu;
      }
    } else {
#pragma unroll
      for (uint32_t k = 0; k < kG; k++) e[k] = (uint32_t)__ldg(p + k);
    }
    bool bad = false;
    uint32_t bad_k = 0;
#pragma unroll
    for (uint32_t k = 0; k < kG; k++) {
      bool b1 = false;
      e[k] = map_one<In, Out, SWAP_IN, SWAP_OUT>(e[k], b1);
      if (b1 && !bad) { bad = true; bad_k = k; }
    }
    if (bad) {
      const unsigned long long k = err_key(head + g * kG + bad_k, kTooLarge);
      best = k < best ? k : best;
    }
    uint32_t o[4];
#pragma unroll
    for (uint32_t j = 0; j < 4u; j++) {
      if (sizeof(Out) == 4) o[j] = e[j];
      else if (sizeof(Out) == 2) o[j] = e[2 * j] | (e[2 * j + 1] << 16);
      else o[j] = (e[4 * j] & 0xFFu) | ((e[4 * j + 1] & 0xFFu) << 8) | ((e[4 * j + 2] & 0xFFu) << 16) | (e[4 * j + 3] << 24);
    }
    stg_stream_v4(gout + g, make_uint4(o[0], o[1], o[2], o[3]));
  }
  for (size_t i = tid; i < head; i += nthreads) element(i);
  for (size_t i = head + ngroups * kG + tid; i < len; i += nthreads) element(i);

  if (sizeof(Out) == 1) {
    best = warp_min_u64(best);
    if ((threadIdx.x & 31u) == 0 && best != kNoError) report_error(scr, best);
  }
  if (grid_last_thread(scr)) {
    bpd::write_result_from_key(res, ld_relaxed_u64(&scr->err_key), len);
    scratch_reset(scr);
  }
}

// ---- reductions over bytes: MODE 0 validate_ascii (first byte >= 0x80 is TOO_LARGE), 1 utf8_length_from_latin1 --
template <int MODE>
__global__ void __launch_bounds__(kBlock) k_scan_latin1(const uint8_t *in, size_t len, Scratch *scr, void *out) {
  __shared__ unsigned long long s_part[kWarps];
  const size_t tid = (size_t)blockIdx.x * kBlock + threadIdx.x, nthreads = (size_t)gridDim.x * kBlock;
  const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  size_t head = (16u - (reinterpret_cast<uintptr_t>(in) & 15u)) & 15u;
  if (head > len) head = len;
  const size_t nvec = (len - head) >> 4;
  const uint4 *vi = reinterpret_cast<const uint4 *>(in + head);
  unsigned long long total = 0, best = kNoError;
  auto word = [&](uint32_t w, size_t i) {
    const uint32_t hi = w & 0x80808080u;
    if (MODE == 0) {
      if (hi) {
        const unsigned long long k = err_key(i + ((__ffs(hi) - 1) >> 3), kTooLarge);
        best = k < best ? k : best;
      }
    } else {
      total += __popc(hi);
    }
  };
  for (size_t v = tid; v < nvec; v += nthreads) {
    const uint4 x = ldg_stream_v4(vi + v);
    const size_t i = head + 16 * v;
    if (MODE == 0) {
      if ((x.x | x.y | x.z | x.w) & 0x80808080u) { word(x.w, i + 12); word(x.z, i + 8); word(x.y, i + 4); word(x.x, i); }
    } else {
      total += __popc(((x.x >> 7) & 0x01010101u) | ((x.y >> 6) & 0x02020202u) | ((x.z >> 5) & 0x04040404u) | ((x.w >> 4) & 0x08080808u));
    }
  }
  for (size_t i = tid; i < head; i += nthreads) word(in[i], i);
  for (size_t i = head + 16 * nvec + tid; i < len; i += nthreads) word(in[i], i);
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
    else *static_cast<unsigned long long *>(out) = ld_relaxed_u64(&scr->acc0) + len;
    scratch_reset(scr);
  }
}

template <class In, class Out, bool SWAP_IN, bool SWAP_OUT>
cudaError_t launch_map(const LaunchCtx &c, const In *in, size_t len, Out *out, void *res) {
  constexpr size_t kG = 16 / sizeof(Out);
  const unsigned long long want = (len / kG + kBlock - 1) / kBlock + 1;
  const unsigned long long cap = (unsigned long long)c.sm_count * 16;
  const unsigned grid = (unsigned)(want < cap ? want : cap);
  k_latin1_map<In, Out, SWAP_IN, SWAP_OUT><<<grid, kBlock, 0, c.stream>>>(in, len, out, c.scratch, static_cast<ResultPOD *>(res));
  count_launch(1);
  return cudaGetLastError();
}

}  // namespace

size_t latin1_family_tiles(const void *in, size_t bytes) { return workspace_slots(tiles_for(in, bytes)); }

cudaError_t launch_convert_latin1_to_utf8(const LaunchCtx &c, const char *in, size_t len, char *out, void *res) {
  return launch_elem<L1ToU8>(c, in, len, out, res);
}
cudaError_t launch_convert_utf8_to_latin1(const LaunchCtx &c, const char *in, size_t len, char *out, void *res) {
  return launch_elem<U8ToL1>(c, in, len, out, res);
}
cudaError_t launch_convert_latin1_to_utf16(const LaunchCtx &c, const char *in, size_t len, uint16_t *out, void *res, bool big_endian) {
  const uint8_t *p = reinterpret_cast<const uint8_t *>(in);
  return big_endian ? launch_map<uint8_t, uint16_t, false, true>(c, p, len, out, res)
                    : launch_map<uint8_t, uint16_t, false, false>(c, p, len, out, res);
}
cudaError_t launch_convert_latin1_to_utf32(const LaunchCtx &c, const char *in, size_t len, uint32_t *out, void *res) {
  return launch_map<uint8_t, uint32_t, false, false>(c, reinterpret_cast<const uint8_t *>(in), len, out, res);
}
cudaError_t launch_convert_utf16_to_latin1(const LaunchCtx &c, const uint16_t *in, size_t len, char *out, void *res, bool big_endian) {
  uint8_t *o = reinterpret_cast<uint8_t *>(out);
  return big_endian ? launch_map<uint16_t, uint8_t, true, false>(c, in, len, o, res)
                    : launch_map<uint16_t, uint8_t, false, false>(c, in, len, o, res);
}
cudaError_t launch_convert_utf32_to_latin1(const LaunchCtx &c, const uint32_t *in, size_t len, char *out, void *res) {
  return launch_map<uint32_t, uint8_t, false, false>(c, in, len, reinterpret_cast<uint8_t *>(out), res);
}
cudaError_t launch_scan_latin1(const LaunchCtx &c, const char *in, size_t len, void *out, int mode) {
  const unsigned long long want = (len / 16 + kBlock - 1) / kBlock + 1;
  const unsigned long long cap = (unsigned long long)c.sm_count * 8;
  const unsigned grid = (unsigned)(want < cap ? want : cap);
  const uint8_t *p = reinterpret_cast<const uint8_t *>(in);
  if (mode == 0) k_scan_latin1<0><<<grid, kBlock, 0, c.stream>>>(p, len, c.scratch, out);
  else k_scan_latin1<1><<<grid, kBlock, 0, c.stream>>>(p, len, c.scratch, out);
  count_launch(1);
  return cudaGetLastError();
}

}  // namespace b200
