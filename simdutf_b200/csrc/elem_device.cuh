// elem_device.cuh — the per-element transcoder shared by the UTF-32 family (k_utf32.cu) and the Latin-1 family
// (k_latin1.cu): ONE launch on the single-pass skeleton (sp_device.cuh) in which every lane owns 64 contiguous input
// bytes, counts its output, hands the warp total to the scan warp, compacts its output into the warp's staging buffer
// and copies the tile out two tiles later, when the look-back has delivered its global offset.
// A conversion is a traits class:
//   In / Out            element types (1, 2 or 4 bytes)
//   kMax                most output elements one input element produces
//   kNeedsNeighbours    whether emit() looks at the element before / after
//   emit(v, prev, next, has_prev, has_next, P, err)  the number of output elements of input element v (what the length
//                       query of the conversion counts), the elements packed little-endian into P (first element
//                       lowest, Out-sized fields) and the element's own error code
//   kFast               byte input only: the trait also has fast_pass1(w, prev, next, bad) — the lane's output count
//                       from SWAR arithmetic on its 16 input words, plus "some element of this lane may be in error"
//                       (conservative) — and emit_word(w, next_word, sp), which stages the output of four input
//                       bytes of a lane that passed the screen; both are used for tiles wholly inside the buffer
// Every element is judged on its own (or with its two neighbours), so the first error is the atomicMin of
// (index << 8 | code).
#pragma once
#include <cstdlib>
#include <type_traits>

#include "bp_device.cuh"
#include "device_common.cuh"
#include "launch.h"
#include "sp_device.cuh"

namespace b200 {
namespace elem {


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

template <class T>
struct Shape {
  using In = typename T::In;
  using Out = typename T::Out;
  static constexpr uint32_t kInPerLane = 64u / sizeof(In);            // 16 code points or 32 units
  static constexpr uint32_t kVec = 16u / sizeof(Out);                 // output elements per 16-byte vector
  static constexpr uint32_t kMaxOut = kInPerLane * T::kMax;           // per lane
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
  if (sizeof(In) == 4) return w[i];
  if (sizeof(In) == 2) return (w[i >> 1] >> (16 * (i & 1))) & 0xFFFFu;
  return (w[i >> 2] >> (8 * (i & 3))) & 0xFFu;
}

// ---- emit pass -------------------------------------------------------------------------------------------
template <class Out>
__device__ __forceinline__ void sts_elem(uint32_t addr, uint32_t v) {
  if (sizeof(Out) == 1) bpd::sts_u8(addr, v);
  else if (sizeof(Out) == 2) bpd::sts_u16(addr, v);
  else bpd::sts_u32(addr, v);
}

// ---- the transcoder: ONE launch on the single-pass skeleton of sp_device.cuh (round 2) ------------------------
// Round 1 ran a counts kernel first (a second read of the input); the lane counts of the emit pass ARE the counts.
// A worker hands its warp total to the scan warp, compacts into the warp's staging buffer at alignment zero (byte
// offset = the lane's exclusive prefix) and copies the tile out two tiles later, when the look-back has delivered its
// global offset (sp::copy_out_bytes: 32-bit words realigned to the destination by one byte permute).
template <class T, int NW>
struct ShapeV3 {
  using S = Shape<T>;
  static constexpr uint32_t kStageBytes = 32u * S::kMaxOut * (uint32_t)sizeof(typename T::Out) + 16u;
  static constexpr uint32_t kSmemBytes = (uint32_t)NW * 2u * kStageBytes;
  static constexpr int kThreads = (NW + 1) * 32;
};

template <class T, int NW, int MINB>
__global__ void __launch_bounds__((NW + 1) * 32, MINB)
k_elem_transcode_v3(const void *ptr, size_t bytes, typename T::Out *out, unsigned long long *desc, uint32_t epoch,
                    uint32_t num_tiles, uint32_t num_cta_tiles, Scratch *scr, ResultPOD *res) {
  using S = Shape<T>;
  using V = ShapeV3<T, NW>;
  using In = typename T::In;
  using Out = typename T::Out;
  extern __shared__ __align__(16) uint32_t smem[];  // [NW][2] staging buffers
  __shared__ sp::Rings rg;
  static_assert(NW <= 31, "one scan warp lane per worker");
  const InView in = make_view_elems(ptr, bytes);
  const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) sp::init_rings(rg, NW);
  __syncthreads();

  if (warp == (unsigned)NW) {
    sp::scan_warp<NW, 1>(rg, desc, epoch, num_cta_tiles, scr, nullptr, nullptr);
  } else {
    const uint32_t stage0 = (uint32_t)__cvta_generic_to_shared(smem) + warp * 2u * V::kStageBytes;
    const long long first_elem = (long long)(in.vbeg / sizeof(In)), end_elem = (long long)(in.vend / sizeof(In));
    uint32_t q1_wtot = 0, q1_iter = 0, q2_wtot = 0, q2_iter = 0;
    bool q1_valid = false, q2_valid = false;
    auto copy_out = [&](uint32_t q_wtot, uint32_t q_iter) {
      const unsigned long long goff = sp::wait_goff(rg, q_iter, warp);
      if (q_wtot)
        sp::copy_out_bytes(stage0 + (q_iter & 1u) * V::kStageBytes, q_wtot * (uint32_t)sizeof(Out),
                           reinterpret_cast<uint8_t *>(out + goff), lane);
      __syncwarp();  // the staging buffer is about to be rewritten
    };

    for (uint32_t iter = 0;; iter++) {
      const uint32_t ct = sp::wait_ticket(rg, iter);
      if (ct >= num_cta_tiles) {  // CTA-uniform: drain
        if (q2_valid) copy_out(q2_wtot, q2_iter);
        if (q1_valid) copy_out(q1_wtot, q1_iter);
        break;
      }
      const uint32_t tile = ct * (uint32_t)NW + warp;
      const uint32_t stage_cur = stage0 + (iter & 1u) * V::kStageBytes;
      const bool active = tile < num_tiles;
      const unsigned long long t0 = (unsigned long long)tile * kTileBytes, r0 = t0 + lane * 64ull;
      const bool interior = active && t0 >= in.vbeg + 16ull && t0 + kTileBytes + 16ull <= in.vend;
      uint32_t w[16];
      if (active) {
        load_lane<In>(in, r0, interior, w);
      } else {
#pragma unroll
        for (int i = 0; i < 16; i++) w[i] = 0u;
      }
      const long long e0 = (long long)(r0 / sizeof(In));  // virtual index of this lane's first element
      uint32_t pv = 0, nv = 0;
      if (T::kNeedsNeighbours && active) {
        pv = elem_guarded<In>(in, e0 - 1);
        nv = elem_guarded<In>(in, e0 + (long long)S::kInPerLane);
      }
      // One element: its output count, packed output and error (filler outside the buffer produces nothing).
      auto eval = [&](int i, uint32_t &Pi, int &err) -> uint32_t {
        const long long idx = e0 + i;
        const bool inside = interior || (active && idx >= first_elem && idx < end_elem);
        const uint32_t v = lane_elem<In>(w, i);
        const uint32_t p = i ? lane_elem<In>(w, i ? i - 1 : 0) : pv;
        const uint32_t nx = i + 1 < (int)S::kInPerLane ? lane_elem<In>(w, i + 1 < (int)S::kInPerLane ? i + 1 : i) : nv;
        uint32_t c = T::emit(v, p, nx, idx > first_elem, idx + 1 < end_elem, Pi, err);
        if (!inside) { c = 0; err = 0; }
        return c;
      };
      auto store = [&](uint32_t &sp_, uint32_t Pi, uint32_t ni) {
#pragma unroll
        for (uint32_t k = 0; k < T::kMax; k++) {
          if (k < ni) sts_elem<Out>(sp_ + k * (uint32_t)sizeof(Out), sizeof(Out) == 4 ? Pi : Pi >> (8u * (uint32_t)sizeof(Out) * k));
        }
        sp_ += ni * (uint32_t)sizeof(Out);
      };
      constexpr bool kKeep = S::kInPerLane <= 32u;
      uint32_t P[kKeep ? S::kInPerLane : 1u];
      uint32_t n[kKeep ? S::kInPerLane : 1u];
      uint32_t cnt = 0;
      long long bad_at = -1;
      int bad_code = 0;
      bool fast = false, suspect = true;
      if constexpr (T::kFast) {
        fast = interior;
        if (fast) cnt = T::fast_pass1(w, pv, nv, suspect);
      }
      if (!fast || suspect) {  // element by element: the count (the same number) and the first error
        cnt = 0;
#pragma unroll
        for (int i = 0; i < (int)S::kInPerLane; i++) {
          int err;
          uint32_t Pi;
          const uint32_t c = eval(i, Pi, err);
          if (err && bad_at < 0) { bad_at = e0 + i; bad_code = err; }
          if (kKeep) { P[i] = Pi; n[i] = c; }
          cnt += c;
        }
      }
      if (bad_at >= 0) {
        const unsigned long long key = err_key((unsigned long long)(bad_at - first_elem), bad_code);
        if (key < ld_relaxed_u64(&scr->err_key)) report_error(scr, key);
      }
      const uint32_t incl = bpd::warp_inclusive_u32(cnt);
      const uint32_t wtot = __shfl_sync(kFull, incl, 31);
      uint32_t tn = 0;
      bool took = sp::post_totals<NW>(rg, iter & 3u, warp, lane, wtot, ct, desc, epoch, scr, tn);
      // ---- the tile before the previous one leaves its staging buffer, which is this tile's ----
      if (q2_valid) copy_out(q2_wtot, q2_iter);
      if (took) {  // the ticket has had the copy-out's time to come back
        tn = sp::post_ticket(rg, iter + 1u, tn, lane);
        if (tn < num_cta_tiles) {  // pull the next CTA-tile into L2
          const char *nx = reinterpret_cast<const char *>(in.base) + (unsigned long long)tn * (NW * kTileBytes);
#pragma unroll
          for (uint32_t k = 0; k < (NW * kTileBytes + 4095u) / 4096u; k++) {
            const uint32_t off = k * 4096u + lane * 128u;
            if (off < NW * kTileBytes) asm volatile("prefetch.global.L2 [%0];" ::"l"(nx + off));
          }
        }
      }
      // ---- compaction into the staging buffer at alignment zero ----
      const uint32_t sp0 = stage_cur + (incl - cnt) * (uint32_t)sizeof(Out);
      bool word_emit = false;
      if constexpr (T::kFast) {
        word_emit = fast && !suspect;  // a screened lane of a tile inside the buffer: four input bytes per step
        if (word_emit) {
          uint32_t sp_ = sp0;
#pragma unroll
          for (int j = 0; j < 16; j++) T::emit_word(w[j], j < 15 ? w[j < 15 ? j + 1 : j] : nv, sp_);
        }
      }
      if (!word_emit) {
        uint32_t sp_ = sp0;
#pragma unroll
        for (int i = 0; i < (int)S::kInPerLane; i++) {
          if (kKeep) {
            store(sp_, P[i], n[i]);
          } else if (fast) {  // inside the buffer: no bounds, no error bookkeeping
            int err;
            uint32_t Pi;
            const uint32_t nx = i + 1 < (int)S::kInPerLane ? lane_elem<In>(w, i + 1 < (int)S::kInPerLane ? i + 1 : i) : nv;
            const uint32_t c = T::emit(lane_elem<In>(w, i), 0u, nx, true, true, Pi, err);
            store(sp_, Pi, c);
          } else {
            int err;
            uint32_t Pi;
            const uint32_t c = eval(i, Pi, err);
            store(sp_, Pi, c);
          }
        }
      }
      __syncwarp();  // the staged elements are visible to the whole warp
      q2_valid = q1_valid; q2_wtot = q1_wtot; q2_iter = q1_iter;
      q1_valid = true; q1_wtot = wtot; q1_iter = iter;
    }
  }

  if (grid_last_thread(scr)) {
    const unsigned long long total = num_cta_tiles ? desc_value(ld_relaxed_u64(desc + (num_cta_tiles - 1u))) : 0ull;
    bpd::write_result_from_key(res, ld_relaxed_u64(&scr->err_key), total);
    scratch_reset(scr);
  }
}

template <class T, int NW, int MINB>
cudaError_t launch_elem_v3(const LaunchCtx &c, const void *in, size_t len, void *out, void *res) {
  using V = ShapeV3<T, NW>;
  const size_t bytes = len * sizeof(typename T::In);
  const size_t span = (reinterpret_cast<uintptr_t>(in) & 15u) + bytes;
  const size_t tiles = (span + kTileBytes - 1) / kTileBytes, cta_tiles = (tiles + NW - 1) / NW;
  if (cta_tiles + 1 > c.desc_capacity || tiles > 0xFFFFFF00ull) return cudaErrorInvalidValue;
  static KernelCache kc;
  int per_sm = 1;
  cudaError_t e = kernel_per_sm(kc, c.device, k_elem_transcode_v3<T, NW, MINB>, V::kThreads, V::kSmemBytes, &per_sm);
  if (e != cudaSuccess) return e;
  const size_t cap = (size_t)c.sm_count * per_sm;
  const unsigned grid = (unsigned)(cta_tiles < cap ? (cta_tiles ? cta_tiles : 1) : cap);
  k_elem_transcode_v3<T, NW, MINB><<<grid, V::kThreads, V::kSmemBytes, c.stream>>>(
      in, bytes, static_cast<typename T::Out *>(out), c.desc, c.epoch, (uint32_t)tiles, (uint32_t)cta_tiles, c.scratch,
      static_cast<ResultPOD *>(res));
  count_launch(1);
  return cudaGetLastError();
}

inline size_t tiles_for(const void *in, size_t bytes) {
  const size_t span = (reinterpret_cast<uintptr_t>(in) & 15u) + bytes;
  return (span + kTileBytes - 1) / kTileBytes;
}
// look-back descriptors (8-byte slots) for CTA-tiles of >= 7 warp-tiles
inline size_t workspace_slots(size_t tiles) { return (tiles + 6) / 7 + 2; }

// Geometry, measured on B200 (1 GiB inputs, ms per launch; workers x CTAs per SM; round 1's two launches first):
//   UTF-32 -> UTF-8   2.016 | 7x4 1.549   8x3 1.543   7x2 2.123   16x1 2.043
//   UTF-32 -> UTF-16  1.363 | 7x4 1.230   8x3 1.184   7x2 1.220
//   UTF-16 -> UTF-32  1.495 | 7x3 1.641   8x3 1.527   7x2 1.573   16x1 1.620   (32 elements per lane keep 64 registers busy)
//   Latin-1 -> UTF-8  1.349 | 7x3 0.946   8x3 0.937   7x2 0.937
//   UTF-8 -> Latin-1  1.644 | 7x4 1.087   8x3 1.103   7x2 1.084
template <class T>
cudaError_t launch_elem(const LaunchCtx &c, const void *in, size_t len, void *out, void *res) {
  return launch_elem_v3<T, 8, 3>(c, in, len, out, res);
}

}  // namespace elem
}  // namespace b200
