// launch.h — host-side launch wrappers implemented next to their kernels (k_*.cu) and used by capi.cu.
// All pointers are DEVICE pointers (or host-mapped pinned memory for result slots).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

namespace b200 {

struct Scratch;

// Everything a launch needs besides its data pointers: the per-stream scratch line, the tile
// descriptor array of the look-back scan (and how many descriptors it holds), the epoch that tags this
// launch's descriptors, the device's SM count, and the stream.
struct LaunchCtx {
  Scratch *scratch;
  unsigned long long *desc;
  size_t desc_capacity;
  unsigned long long *cnt;  // second array of desc_capacity slots: per-tile counts (never aliases descriptors, whose
                            // epoch tags must not be imitated by stale data)
  void *tmp;                // grow-only device scratch of the stream (sized by the caller of get_ws)
  uint32_t epoch;
  int sm_count;
  int device;               // index of the device the stream belongs to (per-device kernel attribute caches)
  cudaStream_t stream;
};

// Per-kernel, per-device cache of "dynamic shared memory attribute set + resident CTAs per SM".  The attribute and the
// occupancy answer belong to a DEVICE, not to the process: one process may drive several devices (b200_set_device,
// the multi-device host path).  Callers hold the device's context mutex, so slot `device` has one writer.
constexpr int kMaxDevices = 16;
struct KernelCache {
  int per_sm[kMaxDevices] = {};
};
template <class Kernel>
inline cudaError_t kernel_per_sm(KernelCache &kc, int device, Kernel kernel, int threads, size_t smem_bytes, int *per_sm) {
  if (device < 0 || device >= kMaxDevices) return cudaErrorInvalidDevice;
  if (kc.per_sm[device] == 0) {
    if (smem_bytes > 48 * 1024) {
      cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
      if (e != cudaSuccess) return e;
    }
    int n = 0;
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kernel, threads, smem_bytes);
    if (e != cudaSuccess) return e;
    kc.per_sm[device] = n < 1 ? 1 : n;
  }
  *per_sm = kc.per_sm[device];
  return cudaSuccess;
}

// Tiles (descriptors) the look-back kernels need for an input of `len` elements.
size_t utf8_to_utf16_tiles(const void *in, size_t len);
size_t utf8_to_utf32_tiles(const void *in, size_t len);
size_t utf16_convert_tiles(const void *in, size_t len);
size_t base64_tiles(const void *in, size_t len);

// result slots: b200_result* / b200_full_result* / uint64_t* (device or mapped-host)
cudaError_t launch_write_result(void *res, int32_t error, unsigned long long count, cudaStream_t stream);
cudaError_t launch_write_full_result(void *res, int32_t error, unsigned long long in_count,
                                     unsigned long long out_count, cudaStream_t stream);
cudaError_t launch_write_u64(unsigned long long *dst, unsigned long long v, cudaStream_t stream);
cudaError_t launch_scratch_init(Scratch *scr, cudaStream_t stream);

cudaError_t launch_validate_utf8(const LaunchCtx &c, const char *in, size_t len, void *res);
cudaError_t launch_count_utf8(const LaunchCtx &c, const char *in, size_t len, unsigned long long *count, int mode);
cudaError_t launch_convert_utf8_to_utf16(const LaunchCtx &c, const char *in, size_t len, uint16_t *out, void *res,
                                         bool big_endian);
cudaError_t launch_convert_utf8_to_utf32(const LaunchCtx &c, const char *in, size_t len, uint32_t *out, void *res);

// UTF-16 operations take the byte order of the units (big_endian = false: UTF-16LE).
cudaError_t launch_count_utf16(const LaunchCtx &c, const uint16_t *in, size_t len, unsigned long long *count, int mode,
                               bool big_endian);
cudaError_t launch_validate_utf16(const LaunchCtx &c, const uint16_t *in, size_t len, void *res, bool big_endian);
cudaError_t launch_convert_utf16_to_utf8(const LaunchCtx &c, const uint16_t *in, size_t len, char *out, void *res,
                                         bool big_endian);
cudaError_t launch_change_endianness_utf16(const LaunchCtx &c, const uint16_t *in, size_t len, uint16_t *out);

cudaError_t launch_base64_to_binary(const LaunchCtx &c, const char *in, size_t len, char *out, uint64_t options,
                                    uint64_t last_chunk, void *full_res);

// to_well_formed_utf16le/be (k_utf16.cu) and the combining step of detect_encodings (k_utf32.cu) — SURVEY.md §8f rank 4
cudaError_t launch_to_well_formed_utf16(const LaunchCtx &c, const uint16_t *in, size_t len, uint16_t *out, bool big_endian);
cudaError_t launch_detect_finish(const LaunchCtx &c, const char *in, size_t len, const void *r8, const void *r16, const void *r32,
                                 unsigned long long *out);
// Latin-1 / ASCII family (k_latin1.cu): scan mode 0 validate_ascii_with_errors (res = b200_result),
// 1 utf8_length_from_latin1 (res = uint64)
size_t latin1_family_tiles(const void *in, size_t bytes);
cudaError_t launch_scan_latin1(const LaunchCtx &c, const char *in, size_t len, void *res, int mode);
cudaError_t launch_convert_latin1_to_utf8(const LaunchCtx &c, const char *in, size_t len, char *out, void *res);
cudaError_t launch_convert_utf8_to_latin1(const LaunchCtx &c, const char *in, size_t len, char *out, void *res);
cudaError_t launch_convert_latin1_to_utf16(const LaunchCtx &c, const char *in, size_t len, uint16_t *out, void *res, bool big_endian);
cudaError_t launch_convert_latin1_to_utf32(const LaunchCtx &c, const char *in, size_t len, uint32_t *out, void *res);
cudaError_t launch_convert_utf16_to_latin1(const LaunchCtx &c, const uint16_t *in, size_t len, char *out, void *res, bool big_endian);
cudaError_t launch_convert_utf32_to_latin1(const LaunchCtx &c, const uint32_t *in, size_t len, char *out, void *res);
// UTF-32 family (k_utf32.cu): scan mode 0 validate_utf32_with_errors (res = b200_result), 1 utf8_length_from_utf32,
// 2 utf16_length_from_utf32 (res = uint64)
size_t utf32_family_tiles(const void *in, size_t bytes);
cudaError_t launch_scan_utf32(const LaunchCtx &c, const uint32_t *in, size_t len, void *res, int mode);
cudaError_t launch_convert_utf32_to_utf8(const LaunchCtx &c, const uint32_t *in, size_t len, char *out, void *res);
cudaError_t launch_convert_utf32_to_utf16(const LaunchCtx &c, const uint32_t *in, size_t len, uint16_t *out, void *res,
                                          bool big_endian);
cudaError_t launch_convert_utf16_to_utf32(const LaunchCtx &c, const uint16_t *in, size_t len, uint32_t *out, void *res,
                                          bool big_endian);

// base64 from char16_t input: narrows into c.tmp (len + 64 bytes), then the byte decoder
cudaError_t launch_base64_to_binary_utf16(const LaunchCtx &c, const uint16_t *in, size_t len, char *out, uint64_t options,
                                          uint64_t last_chunk, void *full_res);
// binary_to_base64: `out` must hold base64_length_from_binary(len, options) characters.
size_t base64_length_from_binary(size_t len, uint64_t options);
cudaError_t launch_binary_to_base64(const LaunchCtx &c, const char *in, size_t len, char *out, uint64_t options);

// combining step of the sharded path (k_sharded.cu): `gathered` = world triplets {input length, result.error, result.count};
// out = b200_sharded_result
cudaError_t launch_sharded_combine(const unsigned long long *gathered, int world, int rank, int count_is_length,
                                   unsigned long long *out, cudaStream_t stream);

// many small strings per launch (k_batch.cu): string i = data[offs[i] .. offs[i + 1]); mode 0 validate (res = b200_result[n]),
// 1 count_utf8, 2 utf16_length_from_utf8 (res = uint64[n])
cudaError_t launch_utf8_batch(int sm_count, cudaStream_t stream, int mode, const char *data, const unsigned long long *offs,
                              unsigned long long n, void *res);
cudaError_t launch_utf8_to_utf16_batch(int sm_count, cudaStream_t stream, bool big_endian, const char *data,
                                       const unsigned long long *offs, unsigned long long n, uint16_t *out,
                                       const unsigned long long *out_offs, void *res);

void count_launch(int n);  // bumps the library-wide launch counter (b200_launch_count)

// Experiment knobs (b200_set_tuning in the C ABI; tools/ and profiles/ use them to compare kernel variants inside
// one build).  The library never reads the environment.
enum TuneKey { kTuneConvMinB = 0, kTuneSegMB = 1, kTuneNoNccl = 2, kTuneConvVariant = 3, kTuneConvStagger = 4, kTuneDbgLo = 5, kTuneDbgHi = 6, kTuneKeyCount = 8 };
int tuning(int key);  // 0 = default (conv_minb and conv_stagger are not read by any kernel at present)

}  // namespace b200
