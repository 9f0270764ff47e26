// bp_device.cuh — small device-side helpers shared by the transcoders: warp sums / scans and the FMA-pipe store helpers
// of the compaction.  (Round 1's "counts" pass — per-tile output counts, chunk offsets by the last CTA — went away when
// every transcoder moved to the single-pass skeleton, sp_device.cuh.)
#pragma once
#include "device_common.cuh"

namespace b200 {
namespace bpd {

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
