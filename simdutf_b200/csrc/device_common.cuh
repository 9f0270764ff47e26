// device_common.cuh — sm_100a building blocks shared by the kernels: guarded 16-byte streaming loads,
// warp reductions, the single-pass decoupled look-back scan over epoch-tagged descriptors, the first-error key,
// and the "last CTA finalises" epilogue.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "swar.h"

namespace b200 {

constexpr int kBlock = 256;          // threads per CTA
constexpr int kWarps = kBlock / 32;  // 8
constexpr unsigned kFull = 0xFFFFFFFFu;

// ---------------------------------------------------------------------------------------------
// Per-stream device scratch.  Its idle state (err_key = ~0, everything else 0) is restored by the
// last CTA of every kernel, so a call is exactly one launch: no memset, no finalise kernel.
// ---------------------------------------------------------------------------------------------
struct Scratch {
  unsigned long long err_key;  // min over (position << 8 | error_code); ~0 = no error
  unsigned long long acc0;     // counters / totals
  unsigned long long acc1;
  unsigned int ticket;         // next tile to hand out
  unsigned int done;           // CTAs finished
  unsigned long long pad[4];
};
static_assert(sizeof(Scratch) == 64, "Scratch is one 64-byte line");

constexpr unsigned long long kNoError = ~0ull;

// ---------------------------------------------------------------------------------------------
// Memory access helpers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 ldg_stream_v4(const uint4 *p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}
// Coherent streaming load: for kernels whose output may alias their input (in-place to_well_formed_utf16 /
// change_endianness_utf16): PTX defines ld.global.nc only for memory that stays read-only for the kernel's lifetime.
__device__ __forceinline__ uint4 ld_cs_v4(const uint4 *p) {
  uint4 r;
  asm volatile("ld.global.cs.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p) : "memory");
  return r;
}
__device__ __forceinline__ void stg_stream_v4(uint4 *p, const uint4 &v) {
  asm volatile("st.global.cs.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ unsigned long long ld_relaxed_u64(const unsigned long long *p) {
  unsigned long long v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_relaxed_u64(unsigned long long *p, unsigned long long v) {
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// The input buffer as seen by the kernels: a 16-byte-aligned base, and the half-open range of
// "virtual" element positions [vbeg, vend) that belong to the caller's buffer (vbeg < 16 bytes'
// worth).  Bytes outside the range read as zero and are never dereferenced.
struct InView {
  const uint4 *base;         // aligned down from the caller's pointer
  unsigned long long vbeg;   // in BYTES from base
  unsigned long long vend;   // in BYTES from base
};

__device__ __forceinline__ uint32_t load_word_guarded(const InView &in, long long word_index) {
  const long long lo = word_index * 4;
  if (lo >= (long long)in.vbeg && lo + 4 <= (long long)in.vend) {
    return __ldg(reinterpret_cast<const uint32_t *>(in.base) + word_index);
  }
  uint32_t r = 0;
  const uint8_t *p = reinterpret_cast<const uint8_t *>(in.base);
#pragma unroll
  for (int b = 0; b < 4; b++) {
    const long long pos = lo + b;
    if (pos >= (long long)in.vbeg && pos < (long long)in.vend) r |= (uint32_t)__ldg(p + pos) << (8 * b);
  }
  return r;
}

// Granule g (16 bytes).  `inside` reports whether all 16 bytes belong to the buffer.
__device__ __forceinline__ void load_granule(const InView &in, unsigned long long g, uint32_t w[4], bool &inside) {
  const unsigned long long lo = g * 16ull;
  inside = lo >= in.vbeg && lo + 16ull <= in.vend;
  if (inside) {
    const uint4 v = ldg_stream_v4(in.base + g);
    w[0] = v.x; w[1] = v.y; w[2] = v.z; w[3] = v.w;
  } else if (lo >= in.vend || lo + 16ull <= in.vbeg) {
    w[0] = w[1] = w[2] = w[3] = 0u;
  } else {
#pragma unroll
    for (int k = 0; k < 4; k++) w[k] = load_word_guarded(in, (long long)(g * 4ull + k));
  }
}

// 16-bit mask (bit b = byte b) of the bytes of granule g that lie inside [vbeg, vend).
__device__ __forceinline__ uint32_t inrange_mask16(const InView &in, unsigned long long g) {
  const long long lo = (long long)(g * 16ull);
  long long a = (long long)in.vbeg - lo, b = (long long)in.vend - lo;
  a = a < 0 ? 0 : (a > 16 ? 16 : a);
  b = b < 0 ? 0 : (b > 16 ? 16 : b);
  if (b <= a) return 0u;
  return ((1u << (unsigned)b) - 1u) & ~((1u << (unsigned)a) - 1u);
}
// Bit-7 mask of the bytes of word `k` of granule g that lie inside [vbeg, vend).
__device__ __forceinline__ uint32_t inrange_mask_word(const InView &in, unsigned long long g, int k) {
  return unmask4((inrange_mask16(in, g) >> (4 * k)) & 0xFu);
}

// ---------------------------------------------------------------------------------------------
// Neighbour words inside a warp-contiguous chunk: lane l, item j owns granule g0 + j*32 + l.
// prev word of (j,l) = last word of granule (j,l-1), or of (j-1,31) for l == 0; symmetric for next.
// The chunk's outer neighbours come from global memory (guarded).
// ---------------------------------------------------------------------------------------------
template <int ITEMS>
__device__ __forceinline__ void neighbour_words(const InView &in, unsigned long long g0, const uint32_t (&w)[ITEMS][4],
                                                uint32_t (&pw)[ITEMS], uint32_t (&nw)[ITEMS]) {
  const unsigned lane = threadIdx.x & 31u;
#pragma unroll
  for (int j = 0; j < ITEMS; j++) {
    // value offered to the lane on my right: my last word (lane 31 offers the last word of item j-1 to lane 0)
    const uint32_t give_p = (lane == 31 && j > 0) ? w[j - 1][3] : w[j][3];
    uint32_t p = __shfl_sync(kFull, give_p, (lane + 31u) & 31u);
    // value offered to the lane on my left: my first word (lane 0 offers the first word of item j+1 to lane 31)
    const uint32_t give_n = (lane == 0 && j < ITEMS - 1) ? w[j + 1][0] : w[j][0];
    uint32_t n = __shfl_sync(kFull, give_n, (lane + 1u) & 31u);
    if (j == 0 && lane == 0) p = load_word_guarded(in, (long long)(g0 * 4ull) - 1);
    if (j == ITEMS - 1 && lane == 31) n = load_word_guarded(in, (long long)((g0 + 32ull * ITEMS) * 4ull));
    pw[j] = p;
    nw[j] = n;
  }
}

// ---------------------------------------------------------------------------------------------
// Warp / block reductions
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long warp_sum_u64(unsigned long long v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
  return v;
}
__device__ __forceinline__ unsigned long long warp_min_u64(unsigned long long v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const unsigned long long t = __shfl_xor_sync(kFull, v, o);
    v = t < v ? t : v;
  }
  return v;
}

// ---------------------------------------------------------------------------------------------
// Single-pass chained scan with decoupled look-back over tile descriptors.
//   descriptor = epoch:12 | status:2 | aux:6 | value:44   (one 64-bit word -> single-copy atomic)
// `epoch` changes every launch (host side), so descriptors never need clearing between launches.
// aux rides along with the nearest contributing tile (base64: the last sextet seen so far).
// Must be called by all 32 lanes of one warp; tiles must be handed out in increasing order by an
// atomic ticket so that every predecessor is resident or finished (forward progress).
// ---------------------------------------------------------------------------------------------
constexpr uint32_t kStatusAggregate = 1, kStatusPrefix = 2;
__device__ __forceinline__ unsigned long long desc_pack(uint32_t epoch, uint32_t status, uint32_t aux,
                                                        unsigned long long value) {
  return ((unsigned long long)epoch << 52) | ((unsigned long long)status << 50) | ((unsigned long long)aux << 44) | value;
}
__device__ __forceinline__ uint32_t desc_epoch(unsigned long long d) { return (uint32_t)(d >> 52); }
__device__ __forceinline__ uint32_t desc_status(unsigned long long d) { return (uint32_t)(d >> 50) & 3u; }
__device__ __forceinline__ uint32_t desc_aux(unsigned long long d) { return (uint32_t)(d >> 44) & 63u; }
__device__ __forceinline__ unsigned long long desc_value(unsigned long long d) { return d & ((1ull << 44) - 1); }

__device__ __forceinline__ void tile_lookback(unsigned long long *desc, uint32_t epoch, uint32_t tile,
                                              unsigned long long agg, uint32_t agg_aux,
                                              unsigned long long &excl, uint32_t &excl_aux) {
  const unsigned lane = threadIdx.x & 31u;
  if (tile == 0) {
    if (lane == 0) st_relaxed_u64(desc, desc_pack(epoch, kStatusPrefix, agg_aux, agg));
    excl = 0;
    excl_aux = 0;
    return;
  }
  if (lane == 0) st_relaxed_u64(desc + tile, desc_pack(epoch, kStatusAggregate, agg_aux, agg));
  unsigned long long sum = 0;
  uint32_t aux = 0;
  bool have_aux = false;
  long long base = (long long)tile - 1;
  while (true) {
    const long long idx = base - (long long)lane;
    unsigned long long d;
    if (idx >= 0) {
      d = ld_relaxed_u64(desc + idx);
      // every predecessor was handed out earlier by the ticket, so this wait is short; the cap only keeps a logic
      // error from hanging the device (the result is then wrong, never silent: callers check the prefix status)
      for (uint32_t spins = 0; (desc_epoch(d) != epoch || desc_status(d) == 0) && spins < (1u << 24); spins++) {
        __nanosleep(32);
        d = ld_relaxed_u64(desc + idx);
      }
    } else {
      d = desc_pack(epoch, kStatusPrefix, 0, 0);  // virtual tiles before the first: empty prefix
    }
    const unsigned pm = __ballot_sync(kFull, desc_status(d) == kStatusPrefix);
    const unsigned first = pm ? (unsigned)(__ffs((int)pm) - 1) : 31u;
    const bool use = lane <= first;
    const unsigned long long v = use ? desc_value(d) : 0ull;
    const unsigned am = __ballot_sync(kFull, use && desc_value(d) != 0ull);
    if (!have_aux && am) {
      aux = __shfl_sync(kFull, desc_aux(d), __ffs((int)am) - 1);
      have_aux = true;
    }
    sum += warp_sum_u64(v);
    if (pm) break;
    base -= 32;
  }
  excl = sum;
  excl_aux = aux;
  if (lane == 0) st_relaxed_u64(desc + tile, desc_pack(epoch, kStatusPrefix, agg ? agg_aux : aux, sum + agg));
}

// ---------------------------------------------------------------------------------------------
// First-error bookkeeping
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long err_key(unsigned long long pos, int code) {
  return (pos << 8) | (unsigned long long)(unsigned)code;
}
__device__ __forceinline__ void report_error(Scratch *scr, unsigned long long key) { atomicMin(&scr->err_key, key); }

// Exact UTF-8 first-error search over byte positions [lo, hi) (virtual positions, clipped to the buffer),
// reading the bytes straight from global memory.  Called only by threads whose granule tripped
// bit-plane detector / the truncated-tail check.  Skips the work if an earlier error is already recorded.  Returns true
// when nothing behind `lo` can be the first error any more (an error was found here, or an earlier one is on record).
static __device__ __noinline__ bool u8_locate_error_impl(const uint4 *base, unsigned long long vbeg,
                                                         unsigned long long vend, Scratch *scr, long long lo,
                                                         long long hi) {
  InView in;
  in.base = base;
  in.vbeg = vbeg;
  in.vend = vend;
  if (lo < (long long)in.vbeg) lo = (long long)in.vbeg;
  if (hi > (long long)in.vend) hi = (long long)in.vend;
  if (lo >= hi) return false;
  const unsigned long long cur = ld_relaxed_u64(&scr->err_key);
  if (cur != kNoError && (cur >> 8) < (unsigned long long)lo - in.vbeg) return true;
  const uint8_t *p = reinterpret_cast<const uint8_t *>(in.base) + in.vbeg;
  const unsigned long long len = in.vend - in.vbeg;
  auto at = [p](unsigned long long j) -> uint32_t { return (uint32_t)p[j]; };
  for (long long v = lo; v < hi; v++) {
    const unsigned long long i = (unsigned long long)v - in.vbeg;
    const int code = u8_verdict(at, i, len);
    if (code != kSuccess) {
      report_error(scr, err_key(i, code));
      return true;
    }
  }
  return false;
}

__device__ __forceinline__ bool u8_locate_error(const InView &in, Scratch *scr, long long lo, long long hi) {
  return u8_locate_error_impl(in.base, in.vbeg, in.vend, scr, lo, hi);
}

// ---------------------------------------------------------------------------------------------
// Last-CTA epilogue: returns true in exactly one thread of the whole grid (thread 0 of the CTA that
// finishes last), after every other CTA's global atomics and stores are visible.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ bool grid_last_thread(Scratch *scr) {
  __threadfence();  // every thread: publish its own global stores before the CTA signs off
  __syncthreads();
  if (threadIdx.x != 0) return false;
  __threadfence();
  const unsigned int d = atomicAdd(&scr->done, 1u);
  if (d != gridDim.x - 1) return false;
  __threadfence();
  return true;
}
__device__ __forceinline__ void scratch_reset(Scratch *scr) {
  scr->err_key = kNoError;
  scr->acc0 = 0;
  scr->acc1 = 0;
  scr->ticket = 0;
  scr->done = 0;
  __threadfence();
}

// POD results as laid out in include/simdutf_b200.h
struct ResultPOD {
  int32_t error;
  uint32_t reserved_;
  unsigned long long count;
};
struct FullResultPOD {
  int32_t error;
  uint32_t reserved_;
  unsigned long long input_count;
  unsigned long long output_count;
};

}  // namespace b200
