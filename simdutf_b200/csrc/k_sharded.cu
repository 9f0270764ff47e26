// k_sharded.cu — the combining step of the sharded (multi-GPU) path, SURVEY.md §8e / BASELINE.json config 5.
//
// A buffer cut on code-point boundaries (simdutf::trim_partial_utf8, reference src/scalar/utf8.h:257-288, used the
// way benchmarks/threaded.cpp:69-88 uses it for two threads) is processed shard by shard with the single-GPU
// kernels.  Every shard leaves one triplet {input length, b200_result{error, count}} in device memory; the triplets
// of all shards are gathered (one NCCL all_gather over NVLink across processes, plain peer/host gathering inside
// one process) and this kernel turns them into what every rank needs, without a host round trip:
//   * the global result{error, count}: the first error in buffer order (minimum of position << 8 | code over the
//     shards — exactly the value an NCCL min-allreduce of that key returns) or {SUCCESS, total output elements};
//   * this shard's global input offset and output offset (exclusive sums), i.e. where its output belongs.
#include "device_common.cuh"
#include "launch.h"

namespace b200 {

namespace {

__global__ void k_sharded_combine(const unsigned long long *gathered, int world, int rank, int count_is_length,
                                  unsigned long long *out) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  unsigned long long in_off = 0, out_off = 0, my_in = 0, my_out = 0, total_in = 0, total_out = 0;
  unsigned long long best = kNoError;
  for (int r = 0; r < world; r++) {
    const unsigned long long n = gathered[3 * r];
    const uint32_t err = (uint32_t)(gathered[3 * r + 1] & 0xFFFFFFFFull);
    const unsigned long long cnt = gathered[3 * r + 2];
    if (r == rank) {
      my_in = in_off;
      my_out = out_off;
    }
    if (err != 0) {
      const unsigned long long key = ((in_off + cnt) << 8) | (unsigned long long)(err & 0xFFu);
      best = key < best ? key : best;
    } else if (!count_is_length) {
      out_off += cnt;
    }
    in_off += n;
  }
  total_in = in_off;
  total_out = out_off;
  if (best == kNoError) {
    out[0] = 0;  // {int32 error = SUCCESS; uint32 reserved}
    out[1] = count_is_length ? total_in : total_out;
  } else {
    out[0] = best & 0xFFull;
    out[1] = best >> 8;
  }
  out[2] = my_in;
  out[3] = my_out;
}

}  // namespace

cudaError_t launch_sharded_combine(const unsigned long long *gathered, int world, int rank, int count_is_length,
                                   unsigned long long *out, cudaStream_t stream) {
  k_sharded_combine<<<1, 32, 0, stream>>>(gathered, world, rank, count_is_length, out);
  count_launch(1);
  return cudaGetLastError();
}

}  // namespace b200
