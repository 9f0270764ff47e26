// capi.cu — the C ABI declared in include/simdutf_b200.h: device contexts, per-stream workspaces, the call flavours
// (async / sync device pointers, host pointers over one or several devices, the sharded multi-device entry points),
// error mapping.
//
// No CPU fallback lives here: every compute entry point ends in a kernel launch from k_*.cu or fails
// with B200_E_NO_DEVICE / a CUDA error code.  The library never prints, throws or aborts
// (reference CMakeLists.txt:173-214 forbids it for anything linked into libsimdutf) and never reads the environment.
//
// State and threading (reference: the library is re-entrant and stateless, include/simdutf/implementation.h:5123-5161;
// benchmarks/threaded.cpp calls it from two threads at once):
//   * DeviceCtx (one per device, process-wide): SM count and the per-STREAM workspaces of caller-provided streams.  Its
//     mutex is held only while a call looks up / grows a workspace and enqueues its launches — never across a
//     synchronisation — so two caller threads overlap their copies and kernels.
//   * HostPath (one per device PER CALLING THREAD): the streams, staging buffers, segment ring and pinned result slots
//     of the host-pointer path.  Nothing in it is shared, so host calls of different threads never wait on each other.
//   * A call never leaves the calling thread's current CUDA device changed (DeviceGuard).
#include <cuda_runtime.h>
#include <dlfcn.h>
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
static std::atomic<int> g_tuning[kTuneKeyCount];
int tuning(int key) { return key >= 0 && key < kTuneKeyCount ? g_tuning[key].load(std::memory_order_relaxed) : 0; }

namespace {

thread_local int tl_device = 0;
thread_local int tl_host_devices = 1;  // devices the host-pointer path of this thread spreads its segments over
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

// Makes `device` current for the lifetime of the guard and restores the caller's device afterwards.
struct DeviceGuard {
  int prev = -1;
  bool changed = false;
  cudaError_t enter(int device) {
    if (cudaGetDevice(&prev) != cudaSuccess) {
      cudaGetLastError();
      prev = -1;
    }
    if (prev != device) {
      const cudaError_t e = cudaSetDevice(device);
      if (e != cudaSuccess) return e;
      changed = prev >= 0;
    }
    return cudaSuccess;
  }
  void leave() {
    if (changed) cudaSetDevice(prev);
    changed = false;
  }
  ~DeviceGuard() { leave(); }
};

// ---------------------------------------------------------------------------------------------
// Per-stream workspace: one Scratch line, the look-back descriptors.
// ---------------------------------------------------------------------------------------------
struct StreamWs {
  Scratch *scratch = nullptr;
  unsigned long long *desc = nullptr;
  unsigned long long *cnt = nullptr;
  size_t desc_cap = 0;
  uint32_t epoch = 0;
  void *tmp = nullptr;     // grow-only device scratch (base64 from char16_t: the narrowed characters)
  size_t tmp_cap = 0;
};

struct DeviceCtx {
  int device = -1;
  int sm_count = 0;
  std::mutex mu;
  std::unordered_map<cudaStream_t, StreamWs> ws;
};

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

// The context of `device`, made current through `guard`; nullptr (and *err set) if unusable.
DeviceCtx *enter_device(int device, DeviceGuard &guard, int *err) {
  const int n = device_count();
  if (n <= 0 || device < 0 || device >= n) {
    *err = fail(B200_E_NO_DEVICE, "no usable sm_100 device");
    return nullptr;
  }
  const cudaError_t e = guard.enter(device);
  if (e != cudaSuccess) {
    *err = fail((int)e, "cudaSetDevice");
    return nullptr;
  }
  *err = 0;
  return &g_ctx[device];
}

// Where does `p` live?  >= 0: device memory (or managed memory) of that device; -1: host memory / unknown to CUDA.
int device_of_pointer(const void *p) {
  if (!p) return -1;
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
    cudaGetLastError();
    return -1;
  }
  if (a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged) return a.device;
  return -1;
}

// Workspace of `stream`, with room for `tiles` descriptors and a fresh epoch.  Caller holds c->mu and has c->device current.
int get_ws(DeviceCtx *c, cudaStream_t stream, size_t tiles, LaunchCtx *lc, size_t tmp_bytes = 0) {
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
  if (!w.scratch) {  // published only once it is allocated AND initialised
    Scratch *s = nullptr;
    B200_CUDA(cudaMalloc(reinterpret_cast<void **>(&s), sizeof(Scratch)));
    const cudaError_t e = launch_scratch_init(s, stream);
    if (e != cudaSuccess) {
      cudaFree(s);
      return fail((int)e, "scratch init");
    }
    w.scratch = s;
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
    unsigned long long *d = nullptr, *k = nullptr;
    B200_CUDA(cudaMalloc(reinterpret_cast<void **>(&d), cap * sizeof(unsigned long long)));
    cudaError_t e = cudaMalloc(reinterpret_cast<void **>(&k), cap * sizeof(unsigned long long));
    if (e == cudaSuccess) e = cudaMemsetAsync(d, 0, cap * sizeof(unsigned long long), stream);
    if (e != cudaSuccess) {
      cudaFree(d);
      if (k) cudaFree(k);
      return fail((int)e, "descriptor allocation");
    }
    w.desc = d;
    w.cnt = k;
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
  return 0;
}

// ---------------------------------------------------------------------------------------------
// HostPath: everything the host-pointer flavour needs on one device, private to the calling thread.
// ---------------------------------------------------------------------------------------------
constexpr int kRing = 3;
constexpr size_t kSmallIn = 32 * 1024;    // inputs up to this many bytes take the zero-copy path
constexpr size_t kSmallOut = 128 * 1024;  // (their output bound is <= 4x the input)

struct HostPath {
  int device = -1;
  bool ok = false;
  cudaStream_t s_main = nullptr, s_copy = nullptr, s_back = nullptr;
  void *h_slot = nullptr;  // 64 B pinned + mapped: kernels write results straight into host memory
  void *d_in = nullptr;    // single-shot staging
  size_t d_in_cap = 0;
  void *d_out = nullptr;
  size_t d_out_cap = 0;
  void *m_in = nullptr;    // zero-copy path: pinned + mapped input / output windows the kernels use directly
  void *m_out = nullptr;
  unsigned long long *d_trip = nullptr;  // sharded entry points: [3] triplet, [3 * kMaxShards] gathered, [4] combined (device)
  void *h_comb = nullptr;                // pinned copy of the combined result
  void *h_pack = nullptr;                // batch entry points: pinned [offsets | packed strings] / results window (grow-only)
  size_t h_pack_cap = 0;
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

  void release() {
    if (device < 0) return;
    DeviceGuard g;
    if (g.enter(device) != cudaSuccess) {
      cudaGetLastError();
      return;
    }
    if (s_main) cudaStreamSynchronize(s_main);
    if (s_copy) cudaStreamSynchronize(s_copy);
    if (s_back) cudaStreamSynchronize(s_back);
    {  // the streams' workspaces
      DeviceCtx &c = g_ctx[device];
      std::lock_guard<std::mutex> lock(c.mu);
      for (cudaStream_t s : {s_main, s_copy, s_back}) {
        auto it = c.ws.find(s);
        if (s && it != c.ws.end()) {
          cudaFree(it->second.scratch);
          cudaFree(it->second.desc);
          cudaFree(it->second.cnt);
          cudaFree(it->second.tmp);
          c.ws.erase(it);
        }
      }
    }
    for (auto &sl : ring) {
      cudaFree(sl.d_in);
      cudaFree(sl.d_out);
      if (sl.h2d_done) cudaEventDestroy(sl.h2d_done);
      if (sl.kernel_done) cudaEventDestroy(sl.kernel_done);
      if (sl.d2h_done) cudaEventDestroy(sl.d2h_done);
      if (sl.h_res) cudaFreeHost(sl.h_res);
      sl = Slot();
    }
    cudaFree(d_in);
    cudaFree(d_out);
    cudaFree(d_trip);
    if (h_slot) cudaFreeHost(h_slot);
    if (h_comb) cudaFreeHost(h_comb);
    if (h_pack) cudaFreeHost(h_pack);
    if (m_in) cudaFreeHost(m_in);
    if (m_out) cudaFreeHost(m_out);
    if (s_main) cudaStreamDestroy(s_main);
    if (s_copy) cudaStreamDestroy(s_copy);
    if (s_back) cudaStreamDestroy(s_back);
    cudaGetLastError();
    *this = HostPath();
  }
};
struct HostPaths {
  HostPath p[kMaxDevices];
  ~HostPaths() {  // thread exit: give the device memory and the streams back
    for (auto &h : p) h.release();
  }
};
thread_local HostPaths tl_paths;

constexpr int kMaxShards = 64;

// This thread's HostPath on the CURRENT device `device` (created on first use).
int host_path(int device, HostPath **out) {
  HostPath &h = tl_paths.p[device];
  if (!h.ok) {
    h.device = device;
    cudaError_t e = cudaStreamCreateWithFlags(&h.s_main, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&h.s_copy, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&h.s_back, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaHostAlloc(&h.h_slot, 64, cudaHostAllocMapped | cudaHostAllocPortable);
    if (e == cudaSuccess) e = cudaHostAlloc(&h.h_comb, 64 * kMaxShards, cudaHostAllocPortable);
    if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void **>(&h.d_trip), 10 * kMaxShards * sizeof(unsigned long long));
    if (e != cudaSuccess) {
      h.release();
      return fail((int)e, "host path init");
    }
    h.ok = true;
  }
  *out = &h;
  return 0;
}

int ensure(void **buf, size_t *cap, size_t need, cudaStream_t s_main) {
  if (need <= *cap) return 0;
  if (*buf) {
    B200_CUDA(cudaStreamSynchronize(s_main));
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

bool bad_args(const void *in, size_t len, const void *res) { return res == nullptr || (in == nullptr && len != 0); }

// Workspace lookup + launches of one operation, atomically with respect to other threads using the same device
// (a multi-launch operation — detect_encodings, base64 from char16_t — must not interleave with another call on the same
// stream's workspace).
int locked_enqueue(DeviceCtx *c, cudaStream_t stream, Op op, const void *in, size_t len, void *out, void *res, uint64_t opt,
                   uint64_t lastc) {
  std::lock_guard<std::mutex> lock(c->mu);
  LaunchCtx lc;
  int err;
  if ((err = get_ws(c, stream, tiles_needed(op, in, len), &lc, tmp_needed(op, len)))) return err;
  return enqueue(op, lc, in, len, out, res, opt, lastc);
}

// Device of a device-pointer call: where the data lives (a tensor on cuda:1 runs on GPU 1 whatever b200_set_device
// said), else the calling thread's selected device.
int call_device(const void *d_in, const void *d_out, const void *d_res) {
  int d = device_of_pointer(d_in);
  if (d < 0) d = device_of_pointer(d_out);
  if (d < 0) d = device_of_pointer(d_res);
  return d >= 0 ? d : tl_device;
}

// async flavour: device pointers, device result slot
int run_async(Op op, const void *d_in, size_t len, void *d_out, void *d_res, void *stream, uint64_t opt = 0, uint64_t lastc = 0) {
  if (bad_args(d_in, len, d_res)) return fail(B200_E_BAD_ARGUMENT, "null pointer");
  int err;
  DeviceGuard guard;
  DeviceCtx *c = enter_device(call_device(d_in, d_out, d_res), guard, &err);
  if (!c) return err;
  return locked_enqueue(c, static_cast<cudaStream_t>(stream), op, d_in, len, d_out, d_res, opt, lastc);
}

// sync flavour: device data pointers, result delivered to the host through this thread's pinned slot
int run_sync_on(int device, cudaStream_t stream, bool own_stream, Op op, const void *d_in, size_t len, void *d_out, void *h_res,
                uint64_t opt, uint64_t lastc) {
  int err;
  DeviceGuard guard;
  DeviceCtx *c = enter_device(device, guard, &err);
  if (!c) return err;
  HostPath *h = nullptr;
  if ((err = host_path(device, &h))) return err;
  if (own_stream) stream = h->s_main;
  if ((err = locked_enqueue(c, stream, op, d_in, len, d_out, h->h_slot, opt, lastc))) return err;
  B200_CUDA(cudaStreamSynchronize(stream));
  std::memcpy(h_res, h->h_slot, result_bytes(op));
  return 0;
}
int run_sync(Op op, const void *d_in, size_t len, void *d_out, void *h_res, void *stream, uint64_t opt = 0, uint64_t lastc = 0) {
  if (bad_args(d_in, len, h_res)) return fail(B200_E_BAD_ARGUMENT, "null pointer");
  return run_sync_on(call_device(d_in, d_out, nullptr), static_cast<cudaStream_t>(stream), false, op, d_in, len, d_out, h_res, opt, lastc);
}

// ---------------------------------------------------------------------------------------------
// Streaming host path, over ONE OR SEVERAL devices.  Large host buffers are cut into segments (on a character boundary
// where the operation validates: the rule of simdutf::trim_partial_utf8 / _utf16le, reference src/scalar/utf8.h:257-288,
// src/scalar/utf16.h:114-124, as benchmarks/threaded.cpp:69-74 uses it) and pipelined through a ring of device staging
// slots: H2D of a later segment (copy stream) overlaps the kernels of the current one (main stream) and the D2H of an
// earlier one (back stream), so a call costs about max(H2D, D2H) instead of H2D + kernel + D2H, and device memory stays
// bounded by the ring whatever the input size.  With n devices (b200_host_set_devices) segment s goes to device s mod n:
// every device has its own ring, streams and PCIe link, so the host-to-host throughput is n links wide; this one
// thread only enqueues asynchronous work and folds results.  Segments are independent calls of the very same kernels;
// their results are combined exactly like the shards of the multi-GPU path (first error in buffer order wins, counts
// add up, a segment's output lands at the sum of the counts before it).
// ---------------------------------------------------------------------------------------------
constexpr size_t kDefaultSegmentMB = 32;
size_t segment_bytes() {
  const int v = tuning(kTuneSegMB);
  return size_t(v >= 1 && v <= 1024 ? v : (int)kDefaultSegmentMB) << 20;
}

// base64 quanta straddle any cut, detect_encodings is three verdicts about the whole buffer: single shot
bool op_streams(Op op) { return op != kOpBase64 && op != kOpBase64U16 && op != kOpDetect; }

// End (exclusive, in elements) of the segment that starts at `beg`.
size_t segment_end(Op op, const void *h_in, size_t len, size_t beg, size_t seg_bytes) {
  const size_t per = seg_bytes / in_elem_bytes(op);
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

int ring_init(HostPath *h) {
  if (h->ring_ok) return 0;
  for (auto &sl : h->ring) {
    B200_CUDA(cudaEventCreateWithFlags(&sl.h2d_done, cudaEventDisableTiming));
    B200_CUDA(cudaEventCreateWithFlags(&sl.kernel_done, cudaEventDisableTiming));
    B200_CUDA(cudaEventCreateWithFlags(&sl.d2h_done, cudaEventDisableTiming));
    B200_CUDA(cudaHostAlloc(&sl.h_res, 64, cudaHostAllocMapped | cudaHostAllocPortable));
  }
  h->ring_ok = true;
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

struct DevSet {
  int n = 0;
  int dev[kMaxDevices];
  HostPath *hp[kMaxDevices];
};

int run_host_streamed(const DevSet &ds, Op op, const void *h_in, size_t len, void *h_out, void *h_res, uint64_t opt,
                      uint64_t lastc) {
  int err;
  const size_t seg_bytes = segment_bytes();
  const size_t ieb = in_elem_bytes(op), oeb = out_elem_bytes(op);
  const bool converts = max_out_bytes(op, 1) != 0;
  const bool is_count = result_bytes(op) == 8;
  const unsigned char *in_bytes = static_cast<const unsigned char *>(h_in);
  unsigned char *out_bytes = static_cast<unsigned char *>(h_out);
  const size_t n = (size_t)ds.n, capacity = n * kRing, window = 2 * n;

  struct Seg { size_t beg, end; };
  Seg segs[kMaxDevices * kRing] = {};
  size_t issued = 0, retired = 0;      // segment counters
  size_t next_beg = 0, out_elems = 0;  // input cursor (elements), output cursor (elements)
  unsigned long long count_sum = 0;
  b200_result final_res = {B200_SUCCESS, 0, 0};
  bool failed = false;
  auto slot_of = [&](size_t s) -> HostPath::Slot & { return ds.hp[s % n]->ring[(s / n) % kRing]; };

  // Everything below runs as a lambda so that ANY failure leaves no copy or kernel in flight on the caller's buffers and
  // no stale `d2h_pending` flag behind (the drain after it).
  auto body = [&]() -> int {
    auto retire = [&](size_t s) -> int {  // wait for segment s, fold its result, start its copy back
      HostPath *h = ds.hp[s % n];
      HostPath::Slot &sl = slot_of(s);
      DeviceGuard g;
      B200_CUDA(g.enter(h->device));
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
          final_res.count = segs[s % capacity].beg + r.count;  // position in the whole buffer
        }
        return 0;
      }
      if (failed) return 0;
      if (converts) {
        if (r.count && out_bytes) {
          B200_CUDA(cudaMemcpyAsync(out_bytes + out_elems * oeb, sl.d_out, (size_t)r.count * oeb, cudaMemcpyDeviceToHost, h->s_back));
          B200_CUDA(cudaEventRecord(sl.d2h_done, h->s_back));
          sl.d2h_pending = true;
        }
        out_elems += (size_t)r.count;
      }
      return 0;
    };

    while (next_beg < len && !failed) {
      // the slot's previous tenant (segment issued - capacity) must be completely done
      while (retired + window <= issued && !failed) {
        if ((err = retire(retired))) return err;
        retired++;
      }
      if (failed) break;
      HostPath *h = ds.hp[issued % n];
      HostPath::Slot &sl = slot_of(issued);
      DeviceGuard g;
      B200_CUDA(g.enter(h->device));
      if (sl.d2h_pending) {
        B200_CUDA(cudaEventSynchronize(sl.d2h_done));
        sl.d2h_pending = false;
      }
      const size_t beg = next_beg, end = segment_end(op, h_in, len, beg, seg_bytes), cnt = end - beg;
      segs[issued % capacity] = {beg, end};
      if ((err = slot_ensure(&sl.d_in, &sl.d_in_cap, seg_bytes + 64))) return err;
      if (converts && (err = slot_ensure(&sl.d_out, &sl.d_out_cap, max_out_bytes(op, seg_bytes / ieb) + 64))) return err;
      B200_CUDA(cudaMemcpyAsync(sl.d_in, in_bytes + beg * ieb, cnt * ieb, cudaMemcpyHostToDevice, h->s_copy));
      B200_CUDA(cudaEventRecord(sl.h2d_done, h->s_copy));
      B200_CUDA(cudaStreamWaitEvent(h->s_main, sl.h2d_done, 0));
      if ((err = locked_enqueue(&g_ctx[h->device], h->s_main, op, sl.d_in, cnt, sl.d_out, sl.h_res, opt, lastc))) return err;
      B200_CUDA(cudaEventRecord(sl.kernel_done, h->s_main));
      issued++;
      next_beg = end;
    }
    while (retired < issued) {
      if ((err = retire(retired))) return err;
      retired++;
    }
    return 0;
  };
  err = body();
  // drain: whatever happened, nothing of this call stays in flight
  for (int i = 0; i < ds.n; i++) {
    HostPath *h = ds.hp[i];
    DeviceGuard g;
    if (g.enter(h->device) != cudaSuccess) continue;
    if (err) {
      cudaStreamSynchronize(h->s_copy);
      cudaStreamSynchronize(h->s_main);
    }
    cudaError_t e = cudaStreamSynchronize(h->s_back);
    for (auto &sl : h->ring) sl.d2h_pending = false;
    if (e != cudaSuccess && !err) err = fail((int)e, "cudaStreamSynchronize(copy back)");
  }
  if (err) {
    cudaGetLastError();
    return err;
  }
  if (is_count) {
    *static_cast<uint64_t *>(h_res) = count_sum;
    return 0;
  }
  if (!failed) final_res.count = converts ? out_elems : len;
  *static_cast<b200_result *>(h_res) = final_res;
  return 0;
}

size_t produced_elems(Op op, const void *h_res, size_t out_cap) {
  size_t produced = 0;
  if (op == kOpBase64 || op == kOpBase64U16) {
    // output_count is 0 on INVALID_BASE64_CHARACTER (unpinned by the reference), so nothing is copied back then
    produced = (size_t)static_cast<const b200_full_result *>(h_res)->output_count;
  } else {
    const b200_result *r = static_cast<const b200_result *>(h_res);
    produced = r->error == B200_SUCCESS ? (size_t)r->count : 0;
  }
  const size_t cap = out_cap / out_elem_bytes(op);
  return produced > cap ? cap : produced;
}

// Small inputs (the reference's test binaries make 10^5..10^8 calls on <= 256-byte strings): no
// staging copies at all.  The text is memcpy'd into a pinned, device-mapped window, the kernels read it and write
// their output and result straight through PCIe into mapped host memory, and one stream synchronisation ends the
// call: launch + sync instead of H2D + launch + sync + D2H + sync.
int run_host_small(HostPath *h, DeviceCtx *c, Op op, const void *h_in, size_t len, void *h_out, void *h_res, uint64_t opt,
                   uint64_t lastc) {
  int err;
  if (!h->m_in) {
    B200_CUDA(cudaHostAlloc(&h->m_in, kSmallIn + 256, cudaHostAllocMapped | cudaHostAllocPortable));
    B200_CUDA(cudaHostAlloc(&h->m_out, kSmallOut + 256, cudaHostAllocMapped | cudaHostAllocPortable));
  }
  const size_t in_bytes = len * in_elem_bytes(op);
  const size_t out_cap = max_out_bytes(op, len);
  // keep the caller's alignment modulo 16: the kernels' paths depend on it and the parity tests sweep it
  char *win = static_cast<char *>(h->m_in) + (reinterpret_cast<uintptr_t>(h_in) & 15u);
  char *wout = static_cast<char *>(h->m_out) + (reinterpret_cast<uintptr_t>(h_out) & 15u);
  std::memcpy(win, h_in, in_bytes);
  if ((err = locked_enqueue(c, h->s_main, op, win, len, wout, h->h_slot, opt, lastc))) return err;
  B200_CUDA(cudaStreamSynchronize(h->s_main));
  std::memcpy(h_res, h->h_slot, result_bytes(op));
  if (out_cap && h_out) {
    const size_t produced = produced_elems(op, h_res, out_cap);
    if (produced) std::memcpy(h_out, wout, produced * out_elem_bytes(op));
  }
  return 0;
}

// host flavour: host pointers in and out (device pointers are recognised and processed in place: the C++ virtuals of
// simdutf::b200::implementation cannot know what their caller holds).
int run_host(Op op, const void *h_in, size_t len, void *h_out, void *h_res, uint64_t opt = 0, uint64_t lastc = 0) {
  if (bad_args(h_in, len, h_res)) return fail(B200_E_BAD_ARGUMENT, "null pointer");
  if (len == 0) {  // never touches CUDA (reference tests/null_safety_tests.cpp:7-95)
    std::memset(h_res, 0, result_bytes(op));
    if (op == kOpDetect) *static_cast<uint64_t *>(h_res) = 1 | 2 | 8;  // "" is valid UTF-8, UTF-16LE and UTF-32LE
    return 0;
  }
  if (device_count() <= 0) return fail(B200_E_NO_DEVICE, "no usable sm_100 device");
  const int pdev = device_of_pointer(h_in);
  if (pdev >= 0) {  // device-resident data: the output pointer must be device memory too
    if (max_out_bytes(op, 1) != 0 && h_out && device_of_pointer(h_out) < 0) return fail(B200_E_BAD_ARGUMENT, "device input with host output");
    return run_sync_on(pdev, nullptr, true, op, h_in, len, h_out, h_res, opt, lastc);
  }
  int err;
  const size_t in_bytes = len * in_elem_bytes(op);
  const size_t out_cap = max_out_bytes(op, len);
  const size_t seg = segment_bytes();
  if (op_streams(op) && in_bytes > seg + seg / 2) {
    DevSet ds;
    const int avail = device_count();
    const int want = tl_host_devices < 1 ? 1 : (tl_host_devices > avail ? avail : tl_host_devices);
    // no point in waking a device for less than two segments of its own
    const size_t nseg = (in_bytes + seg - 1) / seg;
    ds.n = (size_t)want > nseg / 2 ? (int)(nseg / 2 ? nseg / 2 : 1) : want;
    for (int i = 0; i < ds.n; i++) {
      ds.dev[i] = (tl_device + i) % avail;
      DeviceGuard g;
      if (!enter_device(ds.dev[i], g, &err)) return err;
      if ((err = host_path(ds.dev[i], &ds.hp[i]))) return err;
      if ((err = ring_init(ds.hp[i]))) return err;
    }
    return run_host_streamed(ds, op, h_in, len, h_out, h_res, opt, lastc);
  }
  DeviceGuard guard;
  DeviceCtx *c = enter_device(tl_device, guard, &err);
  if (!c) return err;
  HostPath *h = nullptr;
  if ((err = host_path(tl_device, &h))) return err;
  if (in_bytes <= kSmallIn && out_cap <= kSmallOut && tmp_needed(op, len) == 0)
    return run_host_small(h, c, op, h_in, len, h_out, h_res, opt, lastc);
  if ((err = ensure(&h->d_in, &h->d_in_cap, in_bytes + 16, h->s_main))) return err;
  if (out_cap && (err = ensure(&h->d_out, &h->d_out_cap, out_cap + 16, h->s_main))) return err;
  B200_CUDA(cudaMemcpyAsync(h->d_in, h_in, in_bytes, cudaMemcpyHostToDevice, h->s_main));
  if ((err = locked_enqueue(c, h->s_main, op, h->d_in, len, h->d_out, h->h_slot, opt, lastc))) {
    cudaStreamSynchronize(h->s_main);  // the copy reads the caller's buffer: do not return while it is in flight
    return err;
  }
  B200_CUDA(cudaStreamSynchronize(h->s_main));
  std::memcpy(h_res, h->h_slot, result_bytes(op));
  if (out_cap && h_out) {
    const size_t produced = produced_elems(op, h_res, out_cap);
    if (produced) {
      B200_CUDA(cudaMemcpyAsync(h_out, h->d_out, produced * out_elem_bytes(op), cudaMemcpyDeviceToHost, h->s_main));
      B200_CUDA(cudaStreamSynchronize(h->s_main));
    }
  }
  return 0;
}

// ---------------------------------------------------------------------------------------------
// Sharded multi-device entry points (SURVEY.md §8b "_mgpu variants taking per-device shard descriptors", §7 step 5):
// ONE process drives several devices.  Every shard is an independent launch of the single-GPU kernels on its own
// device and stream, leaving the triplet {length, result} in that device's memory; the triplets are exchanged with one
// ncclAllGather over NVLink (communicators from ncclCommInitAll, one rank per device; NCCL is resolved at run time
// with dlopen so that the library has no link-time dependency on it) and k_sharded_combine turns them into the global
// result and the shard's offsets on every device.  Without NCCL, or with several shards on one device, the same
// triplets travel as 24-byte peer copies instead.  Nothing but the final 32-byte results crosses to the host.
// ---------------------------------------------------------------------------------------------
struct NcclApi {
  bool tried = false, ok = false;
  int (*CommInitAll)(void **, int, const int *) = nullptr;
  int (*AllGather)(const void *, void *, size_t, int, void *, cudaStream_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
};
NcclApi g_nccl;
std::mutex g_nccl_mu;
struct CommSet {
  int n = 0;
  int dev[kMaxDevices];
  void *comm[kMaxDevices];
};
CommSet g_comms[8];
int g_comm_sets = 0;

// Communicators for exactly this ordered device list, or nullptr (NCCL absent / init failed).  Caller holds g_nccl_mu.
CommSet *nccl_comms(const int *dev, int n) {
  if (!g_nccl.tried) {
    g_nccl.tried = true;
    void *lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (lib) {
      g_nccl.CommInitAll = reinterpret_cast<int (*)(void **, int, const int *)>(dlsym(lib, "ncclCommInitAll"));
      g_nccl.AllGather = reinterpret_cast<int (*)(const void *, void *, size_t, int, void *, cudaStream_t)>(dlsym(lib, "ncclAllGather"));
      g_nccl.GroupStart = reinterpret_cast<int (*)()>(dlsym(lib, "ncclGroupStart"));
      g_nccl.GroupEnd = reinterpret_cast<int (*)()>(dlsym(lib, "ncclGroupEnd"));
      g_nccl.ok = g_nccl.CommInitAll && g_nccl.AllGather && g_nccl.GroupStart && g_nccl.GroupEnd;
    }
  }
  if (!g_nccl.ok || tuning(kTuneNoNccl)) return nullptr;
  for (int s = 0; s < g_comm_sets; s++) {
    if (g_comms[s].n == n && std::memcmp(g_comms[s].dev, dev, n * sizeof(int)) == 0) return &g_comms[s];
  }
  if (g_comm_sets >= 8) return nullptr;
  CommSet &cs = g_comms[g_comm_sets];
  cs.n = n;
  std::memcpy(cs.dev, dev, n * sizeof(int));
  if (g_nccl.CommInitAll(cs.comm, n, dev) != 0) return nullptr;
  g_comm_sets++;
  return &cs;
}

std::atomic<int> g_last_gather{0};  // 1 = NCCL all_gather, 2 = peer copies (b200_mgpu_last_gather, for tests / logs)

constexpr int kNcclUint64 = 5;  // ncclDataType_t, nccl.h

int run_mgpu(Op op, const b200_shard *shards, int n, b200_sharded_result *h_results, int count_is_length) {
  if (!shards || !h_results || n < 1 || n > kMaxShards) return fail(B200_E_BAD_ARGUMENT, "bad shard list");
  int err;
  const bool is_count = result_bytes(op) == 8;
  HostPath *hp[kMaxShards] = {};
  cudaEvent_t ev[kMaxShards] = {};
  bool distinct = n <= kMaxDevices;
  for (int i = 0; i < n; i++) {
    if (bad_args(shards[i].d_in, shards[i].len, &err)) return fail(B200_E_BAD_ARGUMENT, "null pointer");
    for (int j = 0; j < i; j++) distinct = distinct && shards[i].device != shards[j].device;
  }
  auto cleanup = [&]() {
    for (int i = 0; i < n; i++) {
      if (ev[i]) cudaEventDestroy(ev[i]);
    }
  };
  auto trip_of = [&](int i) { return hp[i]->d_trip + 3 * i; };
  auto gathered_of = [&](int i) { return hp[i]->d_trip + 3 * kMaxShards; };
  auto comb_of = [&](int i) { return hp[i]->d_trip + 6 * kMaxShards + 4 * i; };
  auto body = [&]() -> int {
    // 1. every shard: its triplet header, then the kernels, on its own device and stream
    for (int i = 0; i < n; i++) {
      DeviceGuard g;
      DeviceCtx *c = enter_device(shards[i].device, g, &err);
      if (!c) return err;
      if ((err = host_path(shards[i].device, &hp[i]))) return err;
      HostPath *h = hp[i];
      B200_CUDA(launch_write_u64(trip_of(i), (unsigned long long)shards[i].len, h->s_main));
      if (is_count) B200_CUDA(launch_write_u64(trip_of(i) + 1, 0ull, h->s_main));
      void *res = is_count ? static_cast<void *>(trip_of(i) + 2) : static_cast<void *>(trip_of(i) + 1);
      if ((err = locked_enqueue(c, h->s_main, op, shards[i].d_in, (size_t)shards[i].len, shards[i].d_out, res, 0, 0))) return err;
      B200_CUDA(cudaEventCreateWithFlags(&ev[i], cudaEventDisableTiming));
      B200_CUDA(cudaEventRecord(ev[i], h->s_main));
    }
    // 2. exchange the triplets
    CommSet *cs = nullptr;
    if (distinct) {
      int devs[kMaxDevices];
      for (int i = 0; i < n; i++) devs[i] = shards[i].device;
      std::lock_guard<std::mutex> lock(g_nccl_mu);
      cs = nccl_comms(devs, n);
      if (cs) {
        if (g_nccl.GroupStart() != 0) return fail(B200_E_NO_DEVICE, "ncclGroupStart");
        for (int i = 0; i < n; i++) {
          if (g_nccl.AllGather(trip_of(i), gathered_of(i), 3, kNcclUint64, cs->comm[i], hp[i]->s_main) != 0) {
            g_nccl.GroupEnd();
            return fail(B200_E_NO_DEVICE, "ncclAllGather");
          }
        }
        if (g_nccl.GroupEnd() != 0) return fail(B200_E_NO_DEVICE, "ncclGroupEnd");
      }
    }
    g_last_gather.store(cs ? 1 : 2, std::memory_order_relaxed);
    if (!cs) {
      for (int i = 0; i < n; i++) {
        DeviceGuard g;
        B200_CUDA(g.enter(shards[i].device));
        for (int j = 0; j < n; j++) {
          if (j != i) B200_CUDA(cudaStreamWaitEvent(hp[i]->s_main, ev[j], 0));
          if (shards[i].device == shards[j].device)
            B200_CUDA(cudaMemcpyAsync(gathered_of(i) + 3 * j, hp[j]->d_trip + 3 * j, 24, cudaMemcpyDeviceToDevice, hp[i]->s_main));
          else
            B200_CUDA(cudaMemcpyPeerAsync(gathered_of(i) + 3 * j, shards[i].device, hp[j]->d_trip + 3 * j, shards[j].device, 24, hp[i]->s_main));
        }
      }
    }
    // 3. combine on every device, results to pinned host memory
    for (int i = 0; i < n; i++) {
      DeviceGuard g;
      B200_CUDA(g.enter(shards[i].device));
      B200_CUDA(launch_sharded_combine(gathered_of(i), n, i, count_is_length, comb_of(i), hp[i]->s_main));
      B200_CUDA(cudaMemcpyAsync(static_cast<char *>(hp[i]->h_comb) + 32 * i, comb_of(i), 32, cudaMemcpyDeviceToHost, hp[i]->s_main));
    }
    for (int i = 0; i < n; i++) {
      DeviceGuard g;
      B200_CUDA(g.enter(shards[i].device));
      B200_CUDA(cudaStreamSynchronize(hp[i]->s_main));
      std::memcpy(&h_results[i], static_cast<char *>(hp[i]->h_comb) + 32 * i, sizeof(b200_sharded_result));
    }
    return 0;
  };
  err = body();
  if (err) {
    for (int i = 0; i < n; i++) {
      if (shards[i].device >= 0 && shards[i].device < device_count() && tl_paths.p[shards[i].device].ok) {
        DeviceGuard g;
        if (g.enter(shards[i].device) == cudaSuccess) cudaStreamSynchronize(tl_paths.p[shards[i].device].s_main);
      }
    }
    cudaGetLastError();
  }
  cleanup();
  return err;
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
  DeviceGuard guard;
  if (!enter_device(tl_device, guard, &err)) return err;
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
  DeviceGuard guard;
  if (!enter_device(call_device(d_in, nullptr, nullptr), guard, &err)) return err;
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
  DeviceGuard guard;
  if (!enter_device(call_device(d_gathered, d_out, nullptr), guard, &err)) return err;
  B200_CUDA(launch_sharded_combine(reinterpret_cast<const unsigned long long *>(d_gathered), world, rank, count_is_length,
                                   reinterpret_cast<unsigned long long *>(d_out), static_cast<cudaStream_t>(stream)));
  return 0;
}


// ---- host-pointer path over several devices, experiment knobs ----
int b200_host_set_devices(int n) {
  if (n < 1 || n > device_count()) return fail(B200_E_NO_DEVICE, "b200_host_set_devices: not that many sm_100 devices");
  tl_host_devices = n;
  return 0;
}
int b200_host_get_devices(void) { return tl_host_devices; }
int b200_set_tuning(const char *name, int value) {
  static const char *const names[] = {"conv_minb", "segment_mb", "no_nccl", "conv_variant", "conv_stagger", "dbg_lo", "dbg_hi"};
  for (int k = 0; k < 7; k++) {
    if (name && std::strcmp(name, names[k]) == 0) {
      g_tuning[k].store(value, std::memory_order_relaxed);
      return 0;
    }
  }
  return fail(B200_E_BAD_ARGUMENT, "unknown tuning knob");
}

// ---- many small strings per launch (kernels in k_batch.cu) ----
namespace {
int batch_async(int mode, bool be, const char *d_data, const uint64_t *d_offsets, size_t n, uint16_t *d_out,
                const uint64_t *d_out_offsets, void *d_res, void *stream) {
  if (n == 0) return 0;
  if (!d_data || !d_offsets || !d_res || (mode == 3 && !d_out)) return fail(B200_E_BAD_ARGUMENT, "null pointer");
  int err;
  DeviceGuard guard;
  DeviceCtx *c = enter_device(call_device(d_data, d_offsets, d_res), guard, &err);
  if (!c) return err;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const unsigned long long *offs = reinterpret_cast<const unsigned long long *>(d_offsets);
  if (mode == 3) {
    B200_CUDA(launch_utf8_to_utf16_batch(c->sm_count, s, be, d_data, offs, n, d_out,
                                         reinterpret_cast<const unsigned long long *>(d_out_offsets), d_res));
  } else {
    B200_CUDA(launch_utf8_batch(c->sm_count, s, mode, d_data, offs, n, d_res));
  }
  return 0;
}

// host strings: packed into one pinned window [n + 1 offsets | bytes], one upload, one launch, one download
int batch_host(int mode, const char *const *h_strings, const size_t *h_lens, size_t n, uint16_t *const *h_outs, void *h_res) {
  if (n == 0) return 0;
  if (!h_strings || !h_lens || !h_res || (mode == 3 && !h_outs)) return fail(B200_E_BAD_ARGUMENT, "null pointer");
  if (device_count() <= 0) return fail(B200_E_NO_DEVICE, "no usable sm_100 device");
  size_t total = 0;
  for (size_t i = 0; i < n; i++) {
    if (h_lens[i] && !h_strings[i]) return fail(B200_E_BAD_ARGUMENT, "null string");
    total += h_lens[i];
  }
  int err;
  DeviceGuard guard;
  DeviceCtx *c = enter_device(tl_device, guard, &err);
  if (!c) return err;
  HostPath *h = nullptr;
  if ((err = host_path(tl_device, &h))) return err;
  const size_t off_bytes = (n + 1) * sizeof(unsigned long long);
  const size_t data_off = (off_bytes + 15) & ~size_t(15);
  const size_t in_bytes = data_off + total + 16;
  const size_t res_elem = mode == 0 || mode == 3 ? sizeof(b200_result) : sizeof(uint64_t);
  const size_t res_bytes = (n * res_elem + 15) & ~size_t(15);
  const size_t out_bytes = res_bytes + (mode == 3 ? 2 * total + 16 : 0);
  const size_t pack_need = in_bytes > out_bytes ? in_bytes : out_bytes;
  if (pack_need > h->h_pack_cap) {
    if (h->h_pack) {
      B200_CUDA(cudaStreamSynchronize(h->s_main));
      B200_CUDA(cudaFreeHost(h->h_pack));
      h->h_pack = nullptr;
      h->h_pack_cap = 0;
    }
    const size_t want = pack_need + pack_need / 4 + 4096;
    cudaError_t e = cudaHostAlloc(&h->h_pack, want, cudaHostAllocPortable);
    if (e != cudaSuccess) return fail((int)e, "cudaHostAlloc(batch window)");
    h->h_pack_cap = want;
  }
  if ((err = ensure(&h->d_in, &h->d_in_cap, in_bytes, h->s_main))) return err;
  if ((err = ensure(&h->d_out, &h->d_out_cap, out_bytes + 16, h->s_main))) return err;
  unsigned long long *offs = static_cast<unsigned long long *>(h->h_pack);
  char *bytes = static_cast<char *>(h->h_pack) + data_off;
  size_t at = 0;
  for (size_t i = 0; i < n; i++) {
    offs[i] = at;
    if (h_lens[i]) std::memcpy(bytes + at, h_strings[i], h_lens[i]);
    at += h_lens[i];
  }
  offs[n] = at;
  B200_CUDA(cudaMemcpyAsync(h->d_in, h->h_pack, data_off + total, cudaMemcpyHostToDevice, h->s_main));
  const char *d_data = static_cast<const char *>(h->d_in) + data_off;
  const unsigned long long *d_offs = static_cast<const unsigned long long *>(h->d_in);
  char *d_res = static_cast<char *>(h->d_out);
  uint16_t *d_units = reinterpret_cast<uint16_t *>(d_res + res_bytes);
  if (mode == 3) {
    B200_CUDA(launch_utf8_to_utf16_batch(c->sm_count, h->s_main, false, d_data, d_offs, n, d_units, nullptr, d_res));
  } else {
    B200_CUDA(launch_utf8_batch(c->sm_count, h->s_main, mode, d_data, d_offs, n, d_res));
  }
  // the packed strings are on the device: the window is free to receive the results (and the units)
  B200_CUDA(cudaMemcpyAsync(h->h_pack, h->d_out, out_bytes, cudaMemcpyDeviceToHost, h->s_main));
  B200_CUDA(cudaStreamSynchronize(h->s_main));
  std::memcpy(h_res, h->h_pack, n * res_elem);
  if (mode == 3) {
    const b200_result *r = static_cast<const b200_result *>(h_res);
    const uint16_t *units = reinterpret_cast<const uint16_t *>(static_cast<const char *>(h->h_pack) + res_bytes);
    size_t o = 0;
    for (size_t i = 0; i < n; i++) {
      if (r[i].error == 0 && r[i].count && h_outs[i]) std::memcpy(h_outs[i], units + o, r[i].count * sizeof(uint16_t));
      o += h_lens[i];
    }
  }
  return 0;
}
}  // namespace

int b200_validate_utf8_batch_async(const char *d_data, const uint64_t *d_offsets, size_t n, b200_result *d_results, void *stream) {
  return batch_async(0, false, d_data, d_offsets, n, nullptr, nullptr, d_results, stream);
}
int b200_count_utf8_batch_async(const char *d_data, const uint64_t *d_offsets, size_t n, uint64_t *d_counts, void *stream) {
  return batch_async(1, false, d_data, d_offsets, n, nullptr, nullptr, d_counts, stream);
}
int b200_utf16_length_from_utf8_batch_async(const char *d_data, const uint64_t *d_offsets, size_t n, uint64_t *d_counts, void *stream) {
  return batch_async(2, false, d_data, d_offsets, n, nullptr, nullptr, d_counts, stream);
}
int b200_convert_utf8_to_utf16le_batch_async(const char *d_data, const uint64_t *d_offsets, size_t n, uint16_t *d_out,
                                             const uint64_t *d_out_offsets, b200_result *d_results, void *stream) {
  return batch_async(3, false, d_data, d_offsets, n, d_out, d_out_offsets, d_results, stream);
}
int b200_convert_utf8_to_utf16be_batch_async(const char *d_data, const uint64_t *d_offsets, size_t n, uint16_t *d_out,
                                             const uint64_t *d_out_offsets, b200_result *d_results, void *stream) {
  return batch_async(3, true, d_data, d_offsets, n, d_out, d_out_offsets, d_results, stream);
}
int b200_host_validate_utf8_batch(const char *const *h_strings, const size_t *h_lens, size_t n, b200_result *h_results) {
  return batch_host(0, h_strings, h_lens, n, nullptr, h_results);
}
int b200_host_count_utf8_batch(const char *const *h_strings, const size_t *h_lens, size_t n, uint64_t *h_counts) {
  return batch_host(1, h_strings, h_lens, n, nullptr, h_counts);
}
int b200_host_utf16_length_from_utf8_batch(const char *const *h_strings, const size_t *h_lens, size_t n, uint64_t *h_counts) {
  return batch_host(2, h_strings, h_lens, n, nullptr, h_counts);
}
int b200_host_convert_utf8_to_utf16le_batch(const char *const *h_strings, const size_t *h_lens, size_t n, uint16_t *const *h_outs,
                                            b200_result *h_results) {
  return batch_host(3, h_strings, h_lens, n, h_outs, h_results);
}

// ---- sharded multi-device entry points ----
int b200_mgpu_validate_utf8_with_errors(const b200_shard *shards, int n, b200_sharded_result *h_results) {
  return run_mgpu(kOpValidateUtf8, shards, n, h_results, 1);
}
int b200_mgpu_utf16_length_from_utf8(const b200_shard *shards, int n, b200_sharded_result *h_results) {
  return run_mgpu(kOpUtf16LenFromUtf8, shards, n, h_results, 0);
}
int b200_mgpu_convert_utf8_to_utf16le(const b200_shard *shards, int n, b200_sharded_result *h_results) {
  return run_mgpu(kOpUtf8ToUtf16, shards, n, h_results, 0);
}
int b200_mgpu_convert_utf8_to_utf32(const b200_shard *shards, int n, b200_sharded_result *h_results) {
  return run_mgpu(kOpUtf8ToUtf32, shards, n, h_results, 0);
}
int b200_mgpu_convert_utf16le_to_utf8(const b200_shard *shards, int n, b200_sharded_result *h_results) {
  return run_mgpu(kOpUtf16ToUtf8, shards, n, h_results, 0);
}
int b200_mgpu_last_gather(void) { return g_last_gather.load(std::memory_order_relaxed); }

}  // extern "C"
