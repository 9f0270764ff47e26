// capi.cu — the C ABI declared in include/simdutf_b200.h: per-device context, per-stream workspace,
// the three call flavours (async / sync device pointers, host pointers), error mapping.
//
// No CPU fallback lives here: every compute entry point ends in a kernel launch from k_*.cu or fails
// with B200_E_NO_DEVICE / a CUDA error code.  The library never prints, throws or aborts
// (reference CMakeLists.txt:173-214 forbids it for anything linked into libsimdutf).
#include <cuda_runtime.h>
#include <atomic>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>
#include <unordered_map>

#include "../../include/simdutf_b200.h"
#include "device_common.cuh"
#include "launch.h"

namespace b200 {

static std::atomic<unsigned long long> g_launches{0};
void count_launch(int n) { g_launches.fetch_add((unsigned long long)n, std::memory_order_relaxed); }

namespace {

thread_local int tl_device = 0;
thread_local char tl_error[256] = {0};

int fail(int code, const char *what) {
  if (code > 0) {
    const char *s = cudaGetErrorString((cudaError_t)code);
    std::strncpy(tl_error, what, sizeof(tl_error) - 1);
    const size_t n = std::strlen(tl_error);
    if (n + 3 < sizeof(tl_error)) {
      std::strncat(tl_error, ": ", sizeof(tl_error) - n - 1);
      std::strncat(tl_error, s ? s : "?", sizeof(tl_error) - std::strlen(tl_error) - 1);
    }
  } else {
    std::strncpy(tl_error, what, sizeof(tl_error) - 1);
  }
  tl_error[sizeof(tl_error) - 1] = 0;
  return code;
}
#define B200_CUDA(expr)                                   \
  do {                                                    \
    cudaError_t e_ = (expr);                              \
    if (e_ != cudaSuccess) return fail((int)e_, #expr);   \
  } while (0)

// ---------------------------------------------------------------------------------------------
// Per-stream workspace: one Scratch line, the look-back descriptors, a pinned result slot.
// ---------------------------------------------------------------------------------------------
struct StreamWs {
  Scratch *scratch = nullptr;
  unsigned long long *desc = nullptr;
  unsigned long long *cnt = nullptr;
  size_t desc_cap = 0;
  uint32_t epoch = 0;
  void *h_slot = nullptr;  // 64 B pinned + mapped: kernels write results straight into host memory
  void *tmp = nullptr;     // grow-only device scratch (base64 from char16_t: the narrowed characters)
  size_t tmp_cap = 0;
};

struct DeviceCtx {
  int device = -1;
  int sm_count = 0;
  bool ok = false;
  std::mutex mu;
  std::unordered_map<cudaStream_t, StreamWs> ws;
  // host-pointer path
  cudaStream_t s_main = nullptr, s_copy = nullptr, s_back = nullptr;
  cudaEvent_t ev[2] = {nullptr, nullptr};
  void *d_in = nullptr;
  size_t d_in_cap = 0;
  void *d_out = nullptr;
  size_t d_out_cap = 0;
  // streaming host path: a ring of segment slots (device staging in/out, events, a pinned result slot each)
  static constexpr int kRing = 3;
  struct Slot {
    void *d_in = nullptr;
    size_t d_in_cap = 0;
    void *d_out = nullptr;
    size_t d_out_cap = 0;
    cudaEvent_t h2d_done = nullptr, kernel_done = nullptr, d2h_done = nullptr;
    void *h_res = nullptr;  // 64 B pinned + mapped
    bool d2h_pending = false;
  } ring[kRing];
  bool ring_ok = false;
};

constexpr int kMaxDevices = 16;
DeviceCtx g_ctx[kMaxDevices];
std::once_flag g_count_once;
int g_device_count = 0;

void probe_devices() {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    g_device_count = 0;
    return;
  }
  if (n > kMaxDevices) n = kMaxDevices;
  int usable = 0;
  for (int d = 0; d < n; d++) {
    cudaDeviceProp p;
    if (cudaGetDeviceProperties(&p, d) != cudaSuccess) {
      cudaGetLastError();
      break;
    }
    if (p.major != 10) break;  // this library carries sm_100a SASS only; devices must be a 10.x prefix
    g_ctx[d].device = d;
    g_ctx[d].sm_count = p.multiProcessorCount;
    usable++;
  }
  g_device_count = usable;
}

int device_count() {
  std::call_once(g_count_once, probe_devices);
  return g_device_count;
}

// Returns the context of the calling thread's device, made current; nullptr if unusable.
DeviceCtx *current_ctx(int *err) {
  const int n = device_count();
  if (n <= 0 || tl_device < 0 || tl_device >= n) {
    *err = fail(B200_E_NO_DEVICE, "no usable sm_100 device");
    return nullptr;
  }
  int cur = -1;
  if (cudaGetDevice(&cur) != cudaSuccess || cur != tl_device) {
    cudaError_t e = cudaSetDevice(tl_device);
    if (e != cudaSuccess) {
      *err = fail((int)e, "cudaSetDevice");
      return nullptr;
    }
  }
  DeviceCtx *c = &g_ctx[tl_device];
  std::lock_guard<std::mutex> lock(c->mu);
  if (!c->ok) {
    cudaError_t e = cudaStreamCreateWithFlags(&c->s_main, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&c->s_copy, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&c->s_back, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->ev[0], cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->ev[1], cudaEventDisableTiming);
    if (e != cudaSuccess) {
      *err = fail((int)e, "context init");
      return nullptr;
    }
    c->ok = true;
  }
  *err = 0;
  return c;
}

// Workspace of `stream`, with room for `tiles` descriptors and a fresh epoch.  Caller holds c->mu.
int get_ws(DeviceCtx *c, cudaStream_t stream, size_t tiles, LaunchCtx *lc, StreamWs **out_ws, size_t tmp_bytes = 0) {
  StreamWs &w = c->ws[stream];
  if (tmp_bytes > w.tmp_cap) {
    if (w.tmp) {
      B200_CUDA(cudaStreamSynchronize(stream));
      B200_CUDA(cudaFree(w.tmp));
      w.tmp = nullptr;
      w.tmp_cap = 0;
    }
    const size_t want = tmp_bytes + tmp_bytes / 8 + 4096;
    B200_CUDA(cudaMalloc(&w.tmp, want));
    w.tmp_cap = want;
  }
  lc->tmp = w.tmp;
  if (!w.scratch) {
    B200_CUDA(cudaMalloc(reinterpret_cast<void **>(&w.scratch), sizeof(Scratch)));
    B200_CUDA(launch_scratch_init(w.scratch, stream));
    B200_CUDA(cudaHostAlloc(&w.h_slot, 64, cudaHostAllocMapped | cudaHostAllocPortable));
  }
  if (tiles > w.desc_cap) {
    size_t cap = w.desc_cap ? w.desc_cap : 4096;
    while (cap < tiles) cap *= 2;
    if (w.desc) {
      B200_CUDA(cudaStreamSynchronize(stream));
      B200_CUDA(cudaFree(w.desc));
      B200_CUDA(cudaFree(w.cnt));
      w.desc = nullptr;
      w.cnt = nullptr;
      w.desc_cap = 0;
    }
    B200_CUDA(cudaMalloc(reinterpret_cast<void **>(&w.desc), cap * sizeof(unsigned long long)));
    B200_CUDA(cudaMalloc(reinterpret_cast<void **>(&w.cnt), cap * sizeof(unsigned long long)));
    B200_CUDA(cudaMemsetAsync(w.desc, 0, cap * sizeof(unsigned long long), stream));
    w.desc_cap = cap;
    w.epoch = 0;
  }
  if (tiles > 0) {
    w.epoch += 1;
    if (w.epoch > 4095) {  // 12-bit tag wrapped: forget every old descriptor
      B200_CUDA(cudaMemsetAsync(w.desc, 0, w.desc_cap * sizeof(unsigned long long), stream));
      w.epoch = 1;
    }
  }
  lc->scratch = w.scratch;
  lc->desc = w.desc;
  lc->cnt = w.cnt;
  lc->desc_capacity = w.desc_cap;
  lc->epoch = w.epoch;
  lc->sm_count = c->sm_count;
  lc->device = c->device;
  lc->stream = stream;
  if (out_ws) *out_ws = &w;
  return 0;
}

int ensure(void **buf, size_t *cap, size_t need, cudaStream_t s_main, cudaStream_t s_copy) {
  if (need <= *cap) return 0;
  if (*buf) {
    B200_CUDA(cudaStreamSynchronize(s_main));
    B200_CUDA(cudaStreamSynchronize(s_copy));
    B200_CUDA(cudaFree(*buf));
    *buf = nullptr;
    *cap = 0;
  }
  size_t want = need + need / 8 + 4096;
  cudaError_t e = cudaMalloc(buf, want);
  if (e != cudaSuccess) {
    cudaGetLastError();
    want = need;
    e = cudaMalloc(buf, want);
  }
  if (e != cudaSuccess) return fail((int)e, "cudaMalloc(staging)");
  *cap = want;
  return 0;
}

enum Op {
  kOpValidateUtf8,
  kOpCountUtf8,
  kOpUtf16LenFromUtf8,
  kOpUtf8ToUtf16,
  kOpUtf8ToUtf32,
  kOpCountUtf16,
  kOpUtf8LenFromUtf16,
  kOpValidateUtf16,
  kOpUtf16ToUtf8,
  kOpBase64,
  // UTF-16BE twins and change_endianness_utf16
  kOpUtf8ToUtf16BE,
  kOpCountUtf16BE,
  kOpUtf8LenFromUtf16BE,
  kOpValidateUtf16BE,
  kOpUtf16BEToUtf8,
  kOpSwapUtf16,
  kOpBase64Encode,  // binary_to_base64
  kOpBase64U16,     // base64_to_binary for char16_t input
  // UTF-32 family (second part)
  kOpValidateUtf32,
  kOpUtf8LenFromUtf32,
  kOpUtf16LenFromUtf32,
  kOpUtf32ToUtf8,
  kOpUtf32ToUtf16,
  kOpUtf32ToUtf16BE,
  kOpUtf16ToUtf32,
  kOpUtf16BEToUtf32,
  // Latin-1 / ASCII family
  kOpValidateAscii,
  kOpUtf8LenFromLatin1,
  kOpLatin1ToUtf8,
  kOpLatin1ToUtf16,
  kOpLatin1ToUtf16BE,
  kOpLatin1ToUtf32,
  kOpUtf8ToLatin1,
  kOpUtf16ToLatin1,
  kOpUtf16BEToLatin1,
  kOpUtf32ToLatin1,
  // SURVEY.md §8f rank 4
  kOpWellFormedUtf16,
  kOpWellFormedUtf16BE,
  kOpDetect
};

size_t tmp_needed(Op op, size_t len) { return op == kOpBase64U16 ? len + 64 : op == kOpDetect ? 64 : 0; }

size_t tiles_needed(Op op, const void *in, size_t len) {
  switch (op) {
    case kOpUtf8ToUtf16: case kOpUtf8ToUtf16BE: return utf8_to_utf16_tiles(in, len);
    case kOpUtf8ToUtf32: return utf8_to_utf32_tiles(in, len);
    case kOpUtf16ToUtf8: case kOpUtf16BEToUtf8: return utf16_convert_tiles(in, len);
    case kOpBase64: return base64_tiles(in, len);
    case kOpBase64U16: return base64_tiles(nullptr, len + 16);  // the narrowed copy is 16-byte aligned
    case kOpUtf32ToUtf8: case kOpUtf32ToUtf16: case kOpUtf32ToUtf16BE: return utf32_family_tiles(in, 4 * len);
    case kOpUtf16ToUtf32: case kOpUtf16BEToUtf32: return utf32_family_tiles(in, 2 * len);
    case kOpLatin1ToUtf8: case kOpUtf8ToLatin1: return latin1_family_tiles(in, len);
    default: return 0;
  }
}

// Enqueue one operation on lc.stream.  `res` is a device-visible slot of the operation's result type.
int enqueue(Op op, const LaunchCtx &lc, const void *in, size_t len, void *out, void *res, uint64_t opt, uint64_t lastc) {
  if (len == 0) {  // reference: empty input is SUCCESS / 0 everywhere on the hot path
    switch (op) {
      case kOpCountUtf8: case kOpUtf16LenFromUtf8: case kOpCountUtf16: case kOpUtf8LenFromUtf16:
      case kOpCountUtf16BE: case kOpUtf8LenFromUtf16BE: case kOpUtf8LenFromUtf32: case kOpUtf16LenFromUtf32:
      case kOpUtf8LenFromLatin1:
        B200_CUDA(launch_write_u64(static_cast<unsigned long long *>(res), 0, lc.stream));
        return 0;
      case kOpDetect:  // the empty buffer validates as UTF-8, UTF-16LE and UTF-32LE (reference src/fallback/implementation.cpp:15-31)
        B200_CUDA(launch_write_u64(static_cast<unsigned long long *>(res), 1 | 2 | 8, lc.stream));
        return 0;
      case kOpBase64: case kOpBase64U16:
        B200_CUDA(launch_write_full_result(res, B200_SUCCESS, 0, 0, lc.stream));
        return 0;
      default:
        B200_CUDA(launch_write_result(res, B200_SUCCESS, 0, lc.stream));
        return 0;
    }
  }
  switch (op) {
    case kOpValidateUtf8: B200_CUDA(launch_validate_utf8(lc, static_cast<const char *>(in), len, res)); break;
    case kOpCountUtf8: B200_CUDA(launch_count_utf8(lc, static_cast<const char *>(in), len, static_cast<unsigned long long *>(res), 0)); break;
    case kOpUtf16LenFromUtf8: B200_CUDA(launch_count_utf8(lc, static_cast<const char *>(in), len, static_cast<unsigned long long *>(res), 1)); break;
    case kOpUtf8ToUtf16: B200_CUDA(launch_convert_utf8_to_utf16(lc, static_cast<const char *>(in), len, static_cast<uint16_t *>(out), res, false)); break;
    case kOpUtf8ToUtf16BE: B200_CUDA(launch_convert_utf8_to_utf16(lc, static_cast<const char *>(in), len, static_cast<uint16_t *>(out), res, true)); break;
    case kOpUtf8ToUtf32: B200_CUDA(launch_convert_utf8_to_utf32(lc, static_cast<const char *>(in), len, static_cast<uint32_t *>(out), res)); break;
    case kOpCountUtf16: B200_CUDA(launch_count_utf16(lc, static_cast<const uint16_t *>(in), len, static_cast<unsigned long long *>(res), 0, false)); break;
    case kOpUtf8LenFromUtf16: B200_CUDA(launch_count_utf16(lc, static_cast<const uint16_t *>(in), len, static_cast<unsigned long long *>(res), 1, false)); break;
    case kOpValidateUtf16: B200_CUDA(launch_validate_utf16(lc, static_cast<const uint16_t *>(in), len, res, false)); break;
    case kOpUtf16ToUtf8: B200_CUDA(launch_convert_utf16_to_utf8(lc, static_cast<const uint16_t *>(in), len, static_cast<char *>(out), res, false)); break;
    case kOpCountUtf16BE: B200_CUDA(launch_count_utf16(lc, static_cast<const uint16_t *>(in), len, static_cast<unsigned long long *>(res), 0, true)); break;
    case kOpUtf8LenFromUtf16BE: B200_CUDA(launch_count_utf16(lc, static_cast<const uint16_t *>(in), len, static_cast<unsigned long long *>(res), 1, true)); break;
    case kOpValidateUtf16BE: B200_CUDA(launch_validate_utf16(lc, static_cast<const uint16_t *>(in), len, res, true)); break;
    case kOpUtf16BEToUtf8: B200_CUDA(launch_convert_utf16_to_utf8(lc, static_cast<const uint16_t *>(in), len, static_cast<char *>(out), res, true)); break;
    case kOpSwapUtf16:
      B200_CUDA(launch_change_endianness_utf16(lc, static_cast<const uint16_t *>(in), len, static_cast<uint16_t *>(out)));
      B200_CUDA(launch_write_result(res, B200_SUCCESS, len, lc.stream));
      break;
    case kOpValidateUtf32: B200_CUDA(launch_scan_utf32(lc, static_cast<const uint32_t *>(in), len, res, 0)); break;
    case kOpUtf8LenFromUtf32: B200_CUDA(launch_scan_utf32(lc, static_cast<const uint32_t *>(in), len, res, 1)); break;
    case kOpUtf16LenFromUtf32: B200_CUDA(launch_scan_utf32(lc, static_cast<const uint32_t *>(in), len, res, 2)); break;
    case kOpUtf32ToUtf8: B200_CUDA(launch_convert_utf32_to_utf8(lc, static_cast<const uint32_t *>(in), len, static_cast<char *>(out), res)); break;
    case kOpUtf32ToUtf16: B200_CUDA(launch_convert_utf32_to_utf16(lc, static_cast<const uint32_t *>(in), len, static_cast<uint16_t *>(out), res, false)); break;
    case kOpUtf32ToUtf16BE: B200_CUDA(launch_convert_utf32_to_utf16(lc, static_cast<const uint32_t *>(in), len, static_cast<uint16_t *>(out), res, true)); break;
    case kOpUtf16ToUtf32: B200_CUDA(launch_convert_utf16_to_utf32(lc, static_cast<const uint16_t *>(in), len, static_cast<uint32_t *>(out), res, false)); break;
    case kOpUtf16BEToUtf32: B200_CUDA(launch_convert_utf16_to_utf32(lc, static_cast<const uint16_t *>(in), len, static_cast<uint32_t *>(out), res, true)); break;
    case kOpValidateAscii: B200_CUDA(launch_scan_latin1(lc, static_cast<const char *>(in), len, res, 0)); break;
    case kOpUtf8LenFromLatin1: B200_CUDA(launch_scan_latin1(lc, static_cast<const char *>(in), len, res, 1)); break;
    case kOpLatin1ToUtf8: B200_CUDA(launch_convert_latin1_to_utf8(lc, static_cast<const char *>(in), len, static_cast<char *>(out), res)); break;
    case kOpUtf8ToLatin1: B200_CUDA(launch_convert_utf8_to_latin1(lc, static_cast<const char *>(in), len, static_cast<char *>(out), res)); break;
    case kOpLatin1ToUtf16: B200_CUDA(launch_convert_latin1_to_utf16(lc, static_cast<const char *>(in), len, static_cast<uint16_t *>(out), res, false)); break;
    case kOpLatin1ToUtf16BE: B200_CUDA(launch_convert_latin1_to_utf16(lc, static_cast<const char *>(in), len, static_cast<uint16_t *>(out), res, true)); break;
    case kOpLatin1ToUtf32: B200_CUDA(launch_convert_latin1_to_utf32(lc, static_cast<const char *>(in), len, static_cast<uint32_t *>(out), res)); break;
    case kOpUtf16ToLatin1: B200_CUDA(launch_convert_utf16_to_latin1(lc, static_cast<const uint16_t *>(in), len, static_cast<char *>(out), res, false)); break;
    case kOpUtf16BEToLatin1: B200_CUDA(launch_convert_utf16_to_latin1(lc, static_cast<const uint16_t *>(in), len, static_cast<char *>(out), res, true)); break;
    case kOpUtf32ToLatin1: B200_CUDA(launch_convert_utf32_to_latin1(lc, static_cast<const uint32_t *>(in), len, static_cast<char *>(out), res)); break;
    case kOpWellFormedUtf16: case kOpWellFormedUtf16BE:
      B200_CUDA(launch_to_well_formed_utf16(lc, static_cast<const uint16_t *>(in), len, static_cast<uint16_t *>(out), op == kOpWellFormedUtf16BE));
      B200_CUDA(launch_write_result(res, B200_SUCCESS, len, lc.stream));
      break;
    case kOpDetect: {  // three validations of the same resident buffer, then the combining step
      if (reinterpret_cast<uintptr_t>(in) & 3u) return fail(B200_E_BAD_ARGUMENT, "detect_encodings: device buffer must be 4-byte aligned");
      char *slots = static_cast<char *>(lc.tmp);
      B200_CUDA(launch_validate_utf8(lc, static_cast<const char *>(in), len, slots));
      if (len % 2 == 0) B200_CUDA(launch_validate_utf16(lc, static_cast<const uint16_t *>(in), len / 2, slots + 16, false));
      if (len % 4 == 0) B200_CUDA(launch_scan_utf32(lc, static_cast<const uint32_t *>(in), len / 4, slots + 32, 0));
      B200_CUDA(launch_detect_finish(lc, static_cast<const char *>(in), len, slots, slots + 16, slots + 32, static_cast<unsigned long long *>(res)));
      break;
    }
    case kOpBase64U16:
      B200_CUDA(launch_base64_to_binary_utf16(lc, static_cast<const uint16_t *>(in), len, static_cast<char *>(out), opt, lastc, res));
      break;
    case kOpBase64Encode:
      B200_CUDA(launch_binary_to_base64(lc, static_cast<const char *>(in), len, static_cast<char *>(out), opt));
      B200_CUDA(launch_write_result(res, B200_SUCCESS, base64_length_from_binary(len, opt), lc.stream));
      break;
    case kOpBase64: B200_CUDA(launch_base64_to_binary(lc, static_cast<const char *>(in), len, static_cast<char *>(out), opt, lastc, res)); break;
  }
  return 0;
}

size_t result_bytes(Op op) {
  switch (op) {
    case kOpCountUtf8: case kOpUtf16LenFromUtf8: case kOpCountUtf16: case kOpUtf8LenFromUtf16:
    case kOpCountUtf16BE: case kOpUtf8LenFromUtf16BE: case kOpUtf8LenFromUtf32: case kOpUtf16LenFromUtf32:
    case kOpUtf8LenFromLatin1: case kOpDetect: return 8;
    case kOpBase64: case kOpBase64U16: return sizeof(b200_full_result);
    default: return sizeof(b200_result);
  }
}

bool bad_args(const void *in, size_t len, const void *res) { return res == nullptr || (in == nullptr && len != 0); }

// async flavour: device pointers, device result slot
int run_async(Op op, const void *d_in, size_t len, void *d_out, void *d_res, void *stream, uint64_t opt = 0, uint64_t lastc = 0) {
  if (bad_args(d_in, len, d_res)) return fail(B200_E_BAD_ARGUMENT, "null pointer");
  int err;
  DeviceCtx *c = current_ctx(&err);
  if (!c) return err;
  std::lock_guard<std::mutex> lock(c->mu);
  LaunchCtx lc;
  if ((err = get_ws(c, static_cast<cudaStream_t>(stream), tiles_needed(op, d_in, len), &lc, nullptr, tmp_needed(op, len)))) return err;
  return enqueue(op, lc, d_in, len, d_out, d_res, opt, lastc);
}

// sync flavour: device data pointers, result delivered to the host
int run_sync(Op op, const void *d_in, size_t len, void *d_out, void *h_res, void *stream, uint64_t opt = 0, uint64_t lastc = 0) {
  if (bad_args(d_in, len, h_res)) return fail(B200_E_BAD_ARGUMENT, "null pointer");
  int err;
  DeviceCtx *c = current_ctx(&err);
  if (!c) return err;
  std::lock_guard<std::mutex> lock(c->mu);
  LaunchCtx lc;
  StreamWs *w = nullptr;
  if ((err = get_ws(c, static_cast<cudaStream_t>(stream), tiles_needed(op, d_in, len), &lc, &w, tmp_needed(op, len)))) return err;
  if ((err = enqueue(op, lc, d_in, len, d_out, w->h_slot, opt, lastc))) return err;
  B200_CUDA(cudaStreamSynchronize(lc.stream));
  std::memcpy(h_res, w->h_slot, result_bytes(op));
  return 0;
}

// Output elements an operation can produce at most for `len` input elements (sizes the device staging
// buffer of the host path; the caller's own buffer is only ever written up to the returned count).
size_t max_out_bytes(Op op, size_t len) {
  switch (op) {
    case kOpUtf8ToUtf16: case kOpUtf8ToUtf16BE: return 2 * len;       // <= 1 unit per input byte
    case kOpUtf8ToUtf32: return 4 * len;
    case kOpUtf16ToUtf8: case kOpUtf16BEToUtf8: return 3 * len;       // <= 3 bytes per unit
    case kOpSwapUtf16: case kOpWellFormedUtf16: case kOpWellFormedUtf16BE: return 2 * len;
    case kOpBase64Encode: return (len + 2) / 3 * 4;
    case kOpUtf32ToUtf8: case kOpUtf32ToUtf16: case kOpUtf32ToUtf16BE: return 4 * len;  // <= 4 bytes / 2 units per code point
    case kOpUtf16ToUtf32: case kOpUtf16BEToUtf32: return 4 * len;                          // <= 1 word per unit
    case kOpBase64: case kOpBase64U16: return len / 4 * 3 + 3;
    case kOpLatin1ToUtf8: case kOpLatin1ToUtf16: case kOpLatin1ToUtf16BE: return 2 * len;  // <= 2 bytes / 1 unit per byte
    case kOpLatin1ToUtf32: return 4 * len;
    case kOpUtf8ToLatin1: case kOpUtf16ToLatin1: case kOpUtf16BEToLatin1: case kOpUtf32ToLatin1: return len;
    default: return 0;
  }
}
size_t in_elem_bytes(Op op) {
  switch (op) {
    case kOpCountUtf16: case kOpUtf8LenFromUtf16: case kOpValidateUtf16: case kOpUtf16ToUtf8:
    case kOpCountUtf16BE: case kOpUtf8LenFromUtf16BE: case kOpValidateUtf16BE: case kOpUtf16BEToUtf8: case kOpSwapUtf16:
    case kOpBase64U16: case kOpUtf16ToUtf32: case kOpUtf16BEToUtf32: case kOpUtf16ToLatin1: case kOpUtf16BEToLatin1:
    case kOpWellFormedUtf16: case kOpWellFormedUtf16BE: return 2;
    case kOpValidateUtf32: case kOpUtf8LenFromUtf32: case kOpUtf16LenFromUtf32: case kOpUtf32ToUtf8: case kOpUtf32ToUtf16:
    case kOpUtf32ToUtf16BE: case kOpUtf32ToLatin1: return 4;
    default: return 1;
  }
}
size_t out_elem_bytes(Op op) {
  switch (op) {
    case kOpUtf8ToUtf16: case kOpUtf8ToUtf16BE: case kOpSwapUtf16: case kOpUtf32ToUtf16: case kOpUtf32ToUtf16BE:
    case kOpLatin1ToUtf16: case kOpLatin1ToUtf16BE: case kOpWellFormedUtf16: case kOpWellFormedUtf16BE: return 2;
    case kOpUtf8ToUtf32: case kOpUtf16ToUtf32: case kOpUtf16BEToUtf32: case kOpLatin1ToUtf32: return 4;
    default: return 1;
  }
}

// ---------------------------------------------------------------------------------------------
// Streaming host path.  Large host buffers are cut into segments (on a character boundary where the operation
// validates: the rule of simdutf::trim_partial_utf8 / _utf16le, reference src/scalar/utf8.h:257-288,
// src/scalar/utf16.h:114-124, as benchmarks/threaded.cpp:69-74 uses it) and pipelined through a ring of device
// staging slots: H2D of segment c+1 (copy stream) overlaps the kernels of segment c (main stream) and the D2H of
// segment c-1 (back stream), so a call costs about max(H2D, D2H) instead of H2D + kernel + D2H, and device
// memory stays bounded by the ring whatever the input size.  Segments are independent calls of the very same
// kernels; their results are combined exactly like the shards of the multi-GPU path (first error in buffer
// order wins, counts add up).
// ---------------------------------------------------------------------------------------------
// B200_TUNE_SEG_MB overrides the segment size for experiments (tools/).
static const size_t kSegmentBytes = [] {
  const char *e = getenv("B200_TUNE_SEG_MB");
  const long v = (e && *e) ? atol(e) : 0;
  return size_t(v >= 1 && v <= 1024 ? v : 32) << 20;
}();

// base64 quanta straddle any cut, detect_encodings is three verdicts about the whole buffer: single shot
bool op_streams(Op op) { return op != kOpBase64 && op != kOpBase64U16 && op != kOpDetect; }

// End (exclusive, in elements) of the segment that starts at `beg`.
size_t segment_end(Op op, const void *h_in, size_t len, size_t beg) {
  const size_t per = kSegmentBytes / in_elem_bytes(op);
  size_t cut = beg + per;
  if (cut >= len) return len;
  switch (op) {
    case kOpValidateUtf8: case kOpUtf8ToUtf16: case kOpUtf8ToUtf32: case kOpUtf8ToUtf16BE: case kOpUtf8ToLatin1: {
      const unsigned char *p = static_cast<const unsigned char *>(h_in);
      for (int k = 0; k < 3 && cut > beg + 1 && (p[cut] & 0xC0) == 0x80; k++) cut--;
      return cut;
    }
    case kOpValidateUtf16: case kOpUtf16ToUtf8: case kOpUtf16ToUtf32: case kOpWellFormedUtf16: {
      const uint16_t *p = static_cast<const uint16_t *>(h_in);
      if ((p[cut] & 0xFC00u) == 0xDC00u && (p[cut - 1] & 0xFC00u) == 0xD800u) cut--;
      return cut;
    }
    case kOpValidateUtf16BE: case kOpUtf16BEToUtf8: case kOpUtf16BEToUtf32: case kOpWellFormedUtf16BE: {  // the same rule on byte-swapped units
      const uint16_t *p = static_cast<const uint16_t *>(h_in);
      if ((p[cut] & 0x00FCu) == 0x00DCu && (p[cut - 1] & 0x00FCu) == 0x00D8u) cut--;
      return cut;
    }
    case kOpBase64Encode: return beg + per / 3 * 3;  // whole 3-byte groups: only the last segment pads
    default: return cut;  // the counts are sums over any partition; change_endianness is a map
  }
}

int ring_init(DeviceCtx *c) {
  if (c->ring_ok) return 0;
  for (auto &sl : c->ring) {
    B200_CUDA(cudaEventCreateWithFlags(&sl.h2d_done, cudaEventDisableTiming));
    B200_CUDA(cudaEventCreateWithFlags(&sl.kernel_done, cudaEventDisableTiming));
    B200_CUDA(cudaEventCreateWithFlags(&sl.d2h_done, cudaEventDisableTiming));
    B200_CUDA(cudaHostAlloc(&sl.h_res, 64, cudaHostAllocMapped | cudaHostAllocPortable));
  }
  c->ring_ok = true;
  return 0;
}

int slot_ensure(void **buf, size_t *cap, size_t need) {
  if (need <= *cap) return 0;
  if (*buf) {
    B200_CUDA(cudaFree(*buf));
    *buf = nullptr;
    *cap = 0;
  }
  B200_CUDA(cudaMalloc(buf, need));
  *cap = need;
  return 0;
}

int run_host_streamed(DeviceCtx *c, Op op, const void *h_in, size_t len, void *h_out, void *h_res, uint64_t opt,
                      uint64_t lastc) {
  int err;
  if ((err = ring_init(c))) return err;
  const size_t ieb = in_elem_bytes(op), oeb = out_elem_bytes(op);
  const bool converts = max_out_bytes(op, 1) != 0;
  const bool is_count = result_bytes(op) == 8;
  const unsigned char *in_bytes = static_cast<const unsigned char *>(h_in);
  unsigned char *out_bytes = static_cast<unsigned char *>(h_out);

  struct Seg { size_t beg, end; };
  Seg segs[DeviceCtx::kRing];
  size_t issued = 0, retired = 0;      // segment counters
  size_t next_beg = 0, out_elems = 0;  // input cursor (elements), output cursor (elements)
  unsigned long long count_sum = 0;
  b200_result final_res = {B200_SUCCESS, 0, 0};
  bool failed = false;

  auto retire = [&](size_t s) -> int {  // wait for segment s, fold its result, start its copy back
    DeviceCtx::Slot &sl = c->ring[s % DeviceCtx::kRing];
    B200_CUDA(cudaEventSynchronize(sl.kernel_done));
    if (is_count) {
      count_sum += *static_cast<const uint64_t *>(sl.h_res);
      return 0;
    }
    const b200_result r = *static_cast<const b200_result *>(sl.h_res);
    if (r.error != B200_SUCCESS) {
      if (!failed) {
        failed = true;
        final_res.error = r.error;
        final_res.count = segs[s % DeviceCtx::kRing].beg + r.count;  // position in the whole buffer
      }
      return 0;
    }
    if (failed) return 0;
    if (converts) {
      if (r.count && out_bytes) {
        B200_CUDA(cudaMemcpyAsync(out_bytes + out_elems * oeb, sl.d_out, (size_t)r.count * oeb, cudaMemcpyDeviceToHost, c->s_back));
        B200_CUDA(cudaEventRecord(sl.d2h_done, c->s_back));
        sl.d2h_pending = true;
      }
      out_elems += (size_t)r.count;
    }
    return 0;
  };

  while (next_beg < len && !failed) {
    DeviceCtx::Slot &sl = c->ring[issued % DeviceCtx::kRing];
    if (issued >= (size_t)DeviceCtx::kRing) {  // the slot's previous tenant must be completely done
      if (retired + DeviceCtx::kRing <= issued) {
        if ((err = retire(retired))) return err;
        retired++;
        if (failed) break;
      }
      if (sl.d2h_pending) {
        B200_CUDA(cudaEventSynchronize(sl.d2h_done));
        sl.d2h_pending = false;
      }
    }
    const size_t beg = next_beg, end = segment_end(op, h_in, len, beg), n = end - beg;
    segs[issued % DeviceCtx::kRing] = {beg, end};
    if ((err = slot_ensure(&sl.d_in, &sl.d_in_cap, kSegmentBytes + 64))) return err;
    if (converts && (err = slot_ensure(&sl.d_out, &sl.d_out_cap, max_out_bytes(op, kSegmentBytes / ieb) + 64))) return err;
    B200_CUDA(cudaMemcpyAsync(sl.d_in, in_bytes + beg * ieb, n * ieb, cudaMemcpyHostToDevice, c->s_copy));
    B200_CUDA(cudaEventRecord(sl.h2d_done, c->s_copy));
    B200_CUDA(cudaStreamWaitEvent(c->s_main, sl.h2d_done, 0));
    LaunchCtx lc;
    if ((err = get_ws(c, c->s_main, tiles_needed(op, sl.d_in, n), &lc, nullptr, tmp_needed(op, n)))) return err;
    if ((err = enqueue(op, lc, sl.d_in, n, sl.d_out, sl.h_res, opt, lastc))) return err;
    B200_CUDA(cudaEventRecord(sl.kernel_done, c->s_main));
    issued++;
    next_beg = end;
    // keep one segment in flight behind the one just issued: retire everything older
    while (retired + 1 < issued && !failed) {
      if ((err = retire(retired))) return err;
      retired++;
    }
  }
  while (retired < issued) {
    if ((err = retire(retired))) return err;
    retired++;
  }
  for (auto &sl : c->ring) {
    if (sl.d2h_pending) {
      B200_CUDA(cudaEventSynchronize(sl.d2h_done));
      sl.d2h_pending = false;
    }
  }
  if (is_count) {
    *static_cast<uint64_t *>(h_res) = count_sum;
    return 0;
  }
  if (!failed) final_res.count = converts ? out_elems : len;
  *static_cast<b200_result *>(h_res) = final_res;
  return 0;
}

// host flavour: host pointers in and out.  H2D on the context's stream, the kernel, then a D2H of exactly
// the elements the result says were produced.
int run_host(Op op, const void *h_in, size_t len, void *h_out, void *h_res, uint64_t opt = 0, uint64_t lastc = 0) {
  if (bad_args(h_in, len, h_res)) return fail(B200_E_BAD_ARGUMENT, "null pointer");
  if (len == 0) {  // never touches CUDA (reference tests/null_safety_tests.cpp:7-95)
    std::memset(h_res, 0, result_bytes(op));
    if (op == kOpDetect) *static_cast<uint64_t *>(h_res) = 1 | 2 | 8;  // "" is valid UTF-8, UTF-16LE and UTF-32LE
    return 0;
  }
  int err;
  DeviceCtx *c = current_ctx(&err);
  if (!c) return err;
  std::lock_guard<std::mutex> lock(c->mu);
  if (op_streams(op) && len * in_elem_bytes(op) > kSegmentBytes + kSegmentBytes / 2)
    return run_host_streamed(c, op, h_in, len, h_out, h_res, opt, lastc);
  const size_t in_bytes = len * in_elem_bytes(op);
  const size_t out_cap = max_out_bytes(op, len);
  if ((err = ensure(&c->d_in, &c->d_in_cap, in_bytes + 16, c->s_main, c->s_copy))) return err;
  if (out_cap && (err = ensure(&c->d_out, &c->d_out_cap, out_cap + 16, c->s_main, c->s_copy))) return err;
  LaunchCtx lc;
  StreamWs *w = nullptr;
  if ((err = get_ws(c, c->s_main, tiles_needed(op, c->d_in, len), &lc, &w, tmp_needed(op, len)))) return err;
  B200_CUDA(cudaMemcpyAsync(c->d_in, h_in, in_bytes, cudaMemcpyHostToDevice, c->s_main));
  if ((err = enqueue(op, lc, c->d_in, len, c->d_out, w->h_slot, opt, lastc))) return err;
  B200_CUDA(cudaStreamSynchronize(c->s_main));
  std::memcpy(h_res, w->h_slot, result_bytes(op));
  if (out_cap && h_out) {
    size_t produced = 0;
    if (op == kOpBase64 || op == kOpBase64U16) {
      const b200_full_result *r = static_cast<const b200_full_result *>(h_res);
      produced = (size_t)r->output_count;
      // output_count is 0 on INVALID_BASE64_CHARACTER (unpinned by the reference), so nothing is copied back then
    } else {
      const b200_result *r = static_cast<const b200_result *>(h_res);
      produced = r->error == B200_SUCCESS ? (size_t)r->count : 0;
    }
    if (produced > out_cap / out_elem_bytes(op)) produced = out_cap / out_elem_bytes(op);
    if (produced) {
      B200_CUDA(cudaMemcpyAsync(h_out, c->d_out, produced * out_elem_bytes(op), cudaMemcpyDeviceToHost, c->s_main));
      B200_CUDA(cudaStreamSynchronize(c->s_main));
    }
  }
  return 0;
}

}  // namespace
}  // namespace b200

using namespace b200;

extern "C" {

int b200_device_count(void) { return device_count(); }
int b200_set_device(int device) {
  if (device < 0 || device >= device_count()) return fail(B200_E_NO_DEVICE, "no such sm_100 device");
  tl_device = device;
  return 0;
}
int b200_get_device(void) { return tl_device; }
const char *b200_name(void) { return "b200"; }
const char *b200_description(void) { return "NVIDIA B200 (sm_100a) CUDA kernels"; }
uint64_t b200_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }
const char *b200_last_error(void) { return tl_error; }

int b200_host_alloc(void **ptr, size_t bytes) {
  if (!ptr) return fail(B200_E_BAD_ARGUMENT, "null pointer");
  int err;
  if (!current_ctx(&err)) return err;
  B200_CUDA(cudaHostAlloc(ptr, bytes ? bytes : 1, cudaHostAllocPortable));
  return 0;
}
int b200_host_free(void *ptr) {
  if (!ptr) return 0;
  B200_CUDA(cudaFreeHost(ptr));
  return 0;
}

#define B200_DEFINE_RESULT_OP(NAME, OP, INTYPE, RESTYPE)                                                    \
  int b200_##NAME##_async(const INTYPE *d_in, size_t len, RESTYPE *d_res, void *stream) {                   \
    return run_async(OP, d_in, len, nullptr, d_res, stream);                                               \
  }                                                                                                        \
  int b200_##NAME(const INTYPE *d_in, size_t len, RESTYPE *h_res, void *stream) {                          \
    return run_sync(OP, d_in, len, nullptr, h_res, stream);                                                \
  }                                                                                                        \
  int b200_host_##NAME(const INTYPE *h_in, size_t len, RESTYPE *h_res) {                                   \
    return run_host(OP, h_in, len, nullptr, h_res);                                                        \
  }

B200_DEFINE_RESULT_OP(validate_utf8_with_errors, kOpValidateUtf8, char, b200_result)
B200_DEFINE_RESULT_OP(count_utf8, kOpCountUtf8, char, uint64_t)
B200_DEFINE_RESULT_OP(utf16_length_from_utf8, kOpUtf16LenFromUtf8, char, uint64_t)
B200_DEFINE_RESULT_OP(count_utf16le, kOpCountUtf16, uint16_t, uint64_t)
B200_DEFINE_RESULT_OP(utf8_length_from_utf16le, kOpUtf8LenFromUtf16, uint16_t, uint64_t)
B200_DEFINE_RESULT_OP(validate_utf16le_with_errors, kOpValidateUtf16, uint16_t, b200_result)
B200_DEFINE_RESULT_OP(count_utf16be, kOpCountUtf16BE, uint16_t, uint64_t)
B200_DEFINE_RESULT_OP(utf8_length_from_utf16be, kOpUtf8LenFromUtf16BE, uint16_t, uint64_t)
B200_DEFINE_RESULT_OP(validate_utf16be_with_errors, kOpValidateUtf16BE, uint16_t, b200_result)
B200_DEFINE_RESULT_OP(validate_utf32_with_errors, kOpValidateUtf32, uint32_t, b200_result)
B200_DEFINE_RESULT_OP(utf8_length_from_utf32, kOpUtf8LenFromUtf32, uint32_t, uint64_t)
B200_DEFINE_RESULT_OP(utf16_length_from_utf32, kOpUtf16LenFromUtf32, uint32_t, uint64_t)
B200_DEFINE_RESULT_OP(validate_ascii_with_errors, kOpValidateAscii, char, b200_result)
B200_DEFINE_RESULT_OP(utf8_length_from_latin1, kOpUtf8LenFromLatin1, char, uint64_t)
B200_DEFINE_RESULT_OP(detect_encodings, kOpDetect, char, uint64_t)

#define B200_DEFINE_CONVERT_OP(NAME, OP, INTYPE, OUTTYPE)                                                          \
  int b200_##NAME##_async(const INTYPE *d_in, size_t len, OUTTYPE *d_out, b200_result *d_res, void *stream) {      \
    return run_async(OP, d_in, len, d_out, d_res, stream);                                                         \
  }                                                                                                                \
  int b200_##NAME(const INTYPE *d_in, size_t len, OUTTYPE *d_out, b200_result *h_res, void *stream) {              \
    return run_sync(OP, d_in, len, d_out, h_res, stream);                                                          \
  }                                                                                                                \
  int b200_host_##NAME(const INTYPE *h_in, size_t len, OUTTYPE *h_out, b200_result *h_res) {                       \
    return run_host(OP, h_in, len, h_out, h_res);                                                                  \
  }

B200_DEFINE_CONVERT_OP(convert_utf8_to_utf16le, kOpUtf8ToUtf16, char, uint16_t)
B200_DEFINE_CONVERT_OP(convert_utf8_to_utf32, kOpUtf8ToUtf32, char, uint32_t)
B200_DEFINE_CONVERT_OP(convert_utf16le_to_utf8, kOpUtf16ToUtf8, uint16_t, char)
B200_DEFINE_CONVERT_OP(convert_utf8_to_utf16be, kOpUtf8ToUtf16BE, char, uint16_t)
B200_DEFINE_CONVERT_OP(convert_utf16be_to_utf8, kOpUtf16BEToUtf8, uint16_t, char)
B200_DEFINE_CONVERT_OP(change_endianness_utf16, kOpSwapUtf16, uint16_t, uint16_t)
B200_DEFINE_CONVERT_OP(convert_utf32_to_utf8, kOpUtf32ToUtf8, uint32_t, char)
B200_DEFINE_CONVERT_OP(convert_utf32_to_utf16le, kOpUtf32ToUtf16, uint32_t, uint16_t)
B200_DEFINE_CONVERT_OP(convert_utf32_to_utf16be, kOpUtf32ToUtf16BE, uint32_t, uint16_t)
B200_DEFINE_CONVERT_OP(convert_utf16le_to_utf32, kOpUtf16ToUtf32, uint16_t, uint32_t)
B200_DEFINE_CONVERT_OP(convert_utf16be_to_utf32, kOpUtf16BEToUtf32, uint16_t, uint32_t)
B200_DEFINE_CONVERT_OP(convert_latin1_to_utf8, kOpLatin1ToUtf8, char, char)
B200_DEFINE_CONVERT_OP(convert_latin1_to_utf16le, kOpLatin1ToUtf16, char, uint16_t)
B200_DEFINE_CONVERT_OP(convert_latin1_to_utf16be, kOpLatin1ToUtf16BE, char, uint16_t)
B200_DEFINE_CONVERT_OP(convert_latin1_to_utf32, kOpLatin1ToUtf32, char, uint32_t)
B200_DEFINE_CONVERT_OP(convert_utf8_to_latin1, kOpUtf8ToLatin1, char, char)
B200_DEFINE_CONVERT_OP(convert_utf16le_to_latin1, kOpUtf16ToLatin1, uint16_t, char)
B200_DEFINE_CONVERT_OP(convert_utf16be_to_latin1, kOpUtf16BEToLatin1, uint16_t, char)
B200_DEFINE_CONVERT_OP(convert_utf32_to_latin1, kOpUtf32ToLatin1, uint32_t, char)
B200_DEFINE_CONVERT_OP(to_well_formed_utf16le, kOpWellFormedUtf16, uint16_t, uint16_t)
B200_DEFINE_CONVERT_OP(to_well_formed_utf16be, kOpWellFormedUtf16BE, uint16_t, uint16_t)

static bool b64_options_ok(uint64_t options, uint64_t last_chunk) {
  return (options <= 5 || options == 8 || options == 12) && last_chunk <= 2;
}
int b200_base64_to_binary_utf16_async(const uint16_t *d_in, size_t len, char *d_out, uint64_t options, uint64_t last_chunk,
                                      b200_full_result *d_res, void *stream) {
  if (!b64_options_ok(options, last_chunk)) return fail(B200_E_BAD_ARGUMENT, "bad base64 argument");
  return run_async(kOpBase64U16, d_in, len, d_out, d_res, stream, options, last_chunk);
}
int b200_base64_to_binary_utf16(const uint16_t *d_in, size_t len, char *d_out, uint64_t options, uint64_t last_chunk,
                                b200_full_result *h_res, void *stream) {
  if (!b64_options_ok(options, last_chunk)) return fail(B200_E_BAD_ARGUMENT, "bad base64 argument");
  return run_sync(kOpBase64U16, d_in, len, d_out, h_res, stream, options, last_chunk);
}
int b200_host_base64_to_binary_utf16(const uint16_t *h_in, size_t len, char *h_out, uint64_t options, uint64_t last_chunk,
                                     b200_full_result *h_res) {
  if (!b64_options_ok(options, last_chunk)) return fail(B200_E_BAD_ARGUMENT, "bad base64 argument");
  return run_host(kOpBase64U16, h_in, len, h_out, h_res, options, last_chunk);
}
int b200_binary_to_base64_async(const char *d_in, size_t len, char *d_out, uint64_t options, b200_result *d_res, void *stream) {
  if (options > 3) return fail(B200_E_BAD_ARGUMENT, "bad base64 argument");
  return run_async(kOpBase64Encode, d_in, len, d_out, d_res, stream, options, 0);
}
int b200_binary_to_base64(const char *d_in, size_t len, char *d_out, uint64_t options, b200_result *h_res, void *stream) {
  if (options > 3) return fail(B200_E_BAD_ARGUMENT, "bad base64 argument");
  return run_sync(kOpBase64Encode, d_in, len, d_out, h_res, stream, options, 0);
}
int b200_host_binary_to_base64(const char *h_in, size_t len, char *h_out, uint64_t options, b200_result *h_res) {
  if (options > 3) return fail(B200_E_BAD_ARGUMENT, "bad base64 argument");
  return run_host(kOpBase64Encode, h_in, len, h_out, h_res, options, 0);
}
size_t b200_base64_length_from_binary(size_t len, uint64_t options) { return base64_length_from_binary(len, options); }

int b200_base64_to_binary_async(const char *d_in, size_t len, char *d_out, uint64_t options, uint64_t last_chunk,
                                b200_full_result *d_res, void *stream) {
  if (!b64_options_ok(options, last_chunk)) return fail(B200_E_BAD_ARGUMENT, "bad base64 argument");
  return run_async(kOpBase64, d_in, len, d_out, d_res, stream, options, last_chunk);
}
int b200_base64_to_binary(const char *d_in, size_t len, char *d_out, uint64_t options, uint64_t last_chunk,
                          b200_full_result *h_res, void *stream) {
  if (!b64_options_ok(options, last_chunk)) return fail(B200_E_BAD_ARGUMENT, "bad base64 argument");
  return run_sync(kOpBase64, d_in, len, d_out, h_res, stream, options, last_chunk);
}
int b200_host_base64_to_binary(const char *h_in, size_t len, char *h_out, uint64_t options, uint64_t last_chunk,
                               b200_full_result *h_res) {
  if (!b64_options_ok(options, last_chunk)) return fail(B200_E_BAD_ARGUMENT, "bad base64 argument");
  return run_host(kOpBase64, h_in, len, h_out, h_res, options, last_chunk);
}

// O(1) helpers on host data (no device work; these are not data-parallel paths in the reference either).
size_t b200_host_maximal_binary_length_from_base64(const char *h_in, size_t len) {
  // reference src/scalar/base64.h:493-513
  size_t padding = 0;
  if (len > 0 && h_in[len - 1] == '=') {
    padding++;
    if (len > 1 && h_in[len - 2] == '=') padding++;
  }
  const size_t actual = len - padding;
  if (actual % 4 <= 1) return actual / 4 * 3;
  return actual / 4 * 3 + (actual % 4) - 1;
}
size_t b200_host_trim_partial_utf8(const char *h_in, size_t len) {
  // reference src/scalar/utf8.h:257-288
  const unsigned char *p = reinterpret_cast<const unsigned char *>(h_in);
  if (len >= 1 && p[len - 1] >= 0xC0) return len - 1;
  if (len >= 2 && p[len - 2] >= 0xE0) return len - 2;
  if (len >= 3 && p[len - 3] >= 0xF0) return len - 3;
  return len;
}
size_t b200_host_trim_partial_utf16le(const uint16_t *h_in, size_t len) {
  // reference src/scalar/utf16.h:114-124
  if (len <= 1) return len;
  return len - (((h_in[len - 1] & 0xFC00u) == 0xD800u) ? 1 : 0);
}
int b200_trim_partial_utf8(const char *d_in, size_t len, size_t *h_trimmed, void *stream) {
  if (!h_trimmed || (!d_in && len)) return fail(B200_E_BAD_ARGUMENT, "null pointer");
  int err;
  if (!current_ctx(&err)) return err;
  char tail[3] = {0, 0, 0};
  const size_t n = len < 3 ? len : 3;
  if (n) {
    B200_CUDA(cudaMemcpyAsync(tail + (3 - n), d_in + (len - n), n, cudaMemcpyDeviceToHost, static_cast<cudaStream_t>(stream)));
    B200_CUDA(cudaStreamSynchronize(static_cast<cudaStream_t>(stream)));
  }
  // tail holds the last n bytes right-aligned; run the host rule on them
  *h_trimmed = len - n + b200_host_trim_partial_utf8(tail + (3 - n), n);
  return 0;
}

int b200_sharded_combine_async(const uint64_t *d_gathered, int world, int rank, int count_is_length,
                               b200_sharded_result *d_out, void *stream) {
  if (!d_gathered || !d_out || world < 1 || world > 4096 || rank < 0 || rank >= world) return fail(B200_E_BAD_ARGUMENT, "bad sharded_combine argument");
  int err;
  if (!current_ctx(&err)) return err;
  B200_CUDA(launch_sharded_combine(reinterpret_cast<const unsigned long long *>(d_gathered), world, rank, count_is_length,
                                   reinterpret_cast<unsigned long long *>(d_out), static_cast<cudaStream_t>(stream)));
  return 0;
}

}  // extern "C"
