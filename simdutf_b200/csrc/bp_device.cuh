// bp_device.cuh — device-side plumbing shared by the bit-plane transcoders (k_utf8_to_utf16.cu, k_utf16_to_utf8.cu):
// the "counts" pass skeleton (per-tile output counts -> per-chunk totals -> exclusive chunk offsets, finished by the
// last CTA), the lookup of a tile's output offset, and the FMA-pipe store helpers of the compaction.
#pragma once
#include "device_common.cuh"

namespace b200 {
namespace bpd {

constexpr int kWarpsPerCta = 8;
constexpr int kThreads = kWarpsPerCta * 32;
constexpr uint32_t kChunkTiles = 64;  // warp-tiles per chunk (one chunk total / chunk offset)

// Counts pass.  `count_tile(t)` returns (warp-uniformly) the number of output elements of warp-tile t.
// Writes tile_cnt[t] (u16) and, through the CTA that finishes last, chunk_off[0..num_chunks] = exclusive offsets of
// the chunk totals, chunk_off[num_chunks] = grand total.
template <class CountTile>
__device__ __forceinline__ void counts_pass(CountTile count_tile, uint16_t *tile_cnt, unsigned long long *chunk_off,
                                            uint32_t num_tiles, uint32_t num_chunks, Scratch *scr) {
  __shared__ uint32_t s_tot[kWarpsPerCta];
  __shared__ unsigned long long s_warp_sum[kWarpsPerCta];
  __shared__ unsigned long long s_carry;
  __shared__ bool s_last;
  const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  for (uint32_t chunk = blockIdx.x; chunk < num_chunks; chunk += gridDim.x) {
    uint32_t mine = 0;
    for (uint32_t i = warp; i < kChunkTiles; i += kWarpsPerCta) {
      const uint32_t t = chunk * kChunkTiles + i;
      if (t >= num_tiles) break;
      const uint32_t c = count_tile(t);
      if (lane == 0) tile_cnt[t] = (uint16_t)c;
      mine += c;
    }
    if (lane == 0) s_tot[warp] = mine;
    __syncthreads();
    if (threadIdx.x == 0) {
      uint32_t tot = 0;
#pragma unroll
      for (int k = 0; k < kWarpsPerCta; k++) tot += s_tot[k];
      chunk_off[chunk] = tot;  // turned into an exclusive offset below
    }
    __syncthreads();
  }
  // the CTA that finishes last scans the chunk totals (8 Ki chunks per GiB of input at 2 KiB tiles)
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    s_last = atomicAdd(&scr->done, 1u) == gridDim.x - 1;
    s_carry = 0;
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  for (uint32_t base = 0; base < num_chunks; base += kThreads) {
    const uint32_t i = base + threadIdx.x;
    const unsigned long long v = i < num_chunks ? ld_relaxed_u64(chunk_off + i) : 0ull;
    unsigned long long incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned long long t = __shfl_up_sync(kFull, incl, o);
      if (lane >= (unsigned)o) incl += t;
    }
    if (lane == 31) s_warp_sum[warp] = incl;
    __syncthreads();
    unsigned long long before = s_carry;
#pragma unroll
    for (int k = 0; k < kWarpsPerCta; k++)
      if ((unsigned)k < warp) before += s_warp_sum[k];
    if (i < num_chunks) chunk_off[i] = before + incl - v;
    __syncthreads();
    if (threadIdx.x == kThreads - 1) s_carry = before + incl;
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    chunk_off[num_chunks] = s_carry;  // grand total
    scr->done = 0;
    __threadfence();
  }
}

// Output offset of warp-tile `tile`: its chunk's offset + the counts of the chunk's earlier tiles.  The two parts are
// returned separately so that the loads can be issued early and the warp reduction done later.
__device__ __forceinline__ uint32_t tile_before_partial(const uint16_t *tile_cnt, uint32_t tile) {
  const unsigned lane = threadIdx.x & 31u;
  const uint32_t chunk = tile / kChunkTiles, in_chunk = tile % kChunkTiles;
  const uint16_t *c = tile_cnt + (size_t)chunk * kChunkTiles;
  uint32_t before = 0;
  if (lane < in_chunk) before += c[lane];
  if (lane + 32u < in_chunk) before += c[lane + 32u];
  return before;
}
__device__ __forceinline__ uint32_t warp_sum_u32(uint32_t v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
  return v;
}
__device__ __forceinline__ uint32_t warp_inclusive_u32(uint32_t v) {
  const unsigned lane = threadIdx.x & 31u;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t t = __shfl_up_sync(kFull, v, o);
    if (lane >= (unsigned)o) v += t;
  }
  return v;
}

// Shared stores through a 32-bit shared-space address, and address + n on the FMA pipe (`one` must be a register the
// assembler cannot fold, e.g. blockDim.x >> 8): LOP3/SHF/PRMT/IADD3 share the half-rate ALU pipe that bounds these
// kernels, IMAD issues next to it (tools/ubench/pipes.cu).
__device__ __forceinline__ void sts_u8(uint32_t addr, uint32_t v) {
  asm volatile("st.shared.u8 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ void sts_u16(uint32_t addr, uint32_t v) {
  asm volatile("{ .reg .b16 l, h; mov.b32 {l, h}, %1; st.shared.b16 [%0], l; }" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ void sts_u32(uint32_t addr, uint32_t v) {
  asm volatile("st.shared.b32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
template <int N>
__device__ __forceinline__ uint32_t bump(uint32_t addr, uint32_t one) {
  uint32_t r;
  asm("mad.lo.u32 %0, %1, %3, %2;" : "=r"(r) : "r"(one), "r"(addr), "n"(N));
  return r;
}

__device__ __forceinline__ void write_result_from_key(ResultPOD *res, unsigned long long key, unsigned long long ok_count) {
  if (key == kNoError) {
    res->error = kSuccess;
    res->reserved_ = 0;
    res->count = ok_count;
  } else {
    res->error = (int32_t)(key & 0xFFu);
    res->reserved_ = 0;
    res->count = key >> 8;
  }
}

}  // namespace bpd
}  // namespace b200
