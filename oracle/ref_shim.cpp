// oracle/ref_shim.cpp — TEST INFRASTRUCTURE ONLY (never linked into the product).
//
// A thin extern "C" window onto the UNMODIFIED reference library so that Python
// tests / golden-vector generation / the bench's CPU baseline can call the
// reference's own kernels (icelake / haswell / westmere / fallback) through
// ctypes.  The reference sources are compiled where they lie under
// /root/reference (see oracle/Makefile); nothing is copied into this repo.
//
// Every function takes the implementation NAME ("icelake", "haswell",
// "fallback", or "best" for the first runtime-supported entry), looked up via
// simdutf::get_available_implementations() (reference
// include/simdutf/implementation.h:5074-5121, 5170).
#include "simdutf.h"
#include <cstring>
#include <cstdint>
#include <thread>
#include <vector>
#include <atomic>
#include <algorithm>

namespace {
const simdutf::implementation *pick(const char *name) {
  if (name == nullptr || std::strcmp(name, "best") == 0) {
    for (const simdutf::implementation *impl :
         simdutf::get_available_implementations()) {
      if (impl->supported_by_runtime_system()) return impl;
    }
    return nullptr;
  }
  const simdutf::implementation *impl =
      simdutf::get_available_implementations()[name];
  if (impl && !impl->supported_by_runtime_system()) return nullptr;
  return impl;
}
} // namespace

extern "C" {

struct ref_result { int32_t error; uint64_t count; };
struct ref_full_result { int32_t error; uint64_t input_count; uint64_t output_count; };

int ref_has_impl(const char *name) { return pick(name) != nullptr; }

const char *ref_best_name() {
  const simdutf::implementation *impl = pick("best");
  static std::string s;
  s = impl ? impl->name() : std::string("none");
  return s.c_str();
}

int ref_validate_utf8_with_errors(const char *impl, const char *in, size_t len, ref_result *out) {
  auto *i = pick(impl); if (!i) return -1;
  simdutf::result r = i->validate_utf8_with_errors(in, len);
  out->error = int32_t(r.error); out->count = r.count; return 0;
}
int ref_validate_utf8(const char *impl, const char *in, size_t len) {
  auto *i = pick(impl); if (!i) return -1;
  return i->validate_utf8(in, len) ? 1 : 0;
}
int64_t ref_count_utf8(const char *impl, const char *in, size_t len) {
  auto *i = pick(impl); if (!i) return -1;
  return int64_t(i->count_utf8(in, len));
}
int64_t ref_utf16_length_from_utf8(const char *impl, const char *in, size_t len) {
  auto *i = pick(impl); if (!i) return -1;
  return int64_t(i->utf16_length_from_utf8(in, len));
}
int64_t ref_utf32_length_from_utf8(const char *impl, const char *in, size_t len) {
  auto *i = pick(impl); if (!i) return -1;
  return int64_t(i->utf32_length_from_utf8(in, len));
}
int ref_convert_utf8_to_utf16le_with_errors(const char *impl, const char *in, size_t len, char16_t *dst, ref_result *out) {
  auto *i = pick(impl); if (!i) return -1;
  simdutf::result r = i->convert_utf8_to_utf16le_with_errors(in, len, dst);
  out->error = int32_t(r.error); out->count = r.count; return 0;
}
int64_t ref_convert_utf8_to_utf16le(const char *impl, const char *in, size_t len, char16_t *dst) {
  auto *i = pick(impl); if (!i) return -1;
  return int64_t(i->convert_utf8_to_utf16le(in, len, dst));
}
int64_t ref_convert_valid_utf8_to_utf16le(const char *impl, const char *in, size_t len, char16_t *dst) {
  auto *i = pick(impl); if (!i) return -1;
  return int64_t(i->convert_valid_utf8_to_utf16le(in, len, dst));
}
int ref_convert_utf8_to_utf32_with_errors(const char *impl, const char *in, size_t len, char32_t *dst, ref_result *out) {
  auto *i = pick(impl); if (!i) return -1;
  simdutf::result r = i->convert_utf8_to_utf32_with_errors(in, len, dst);
  out->error = int32_t(r.error); out->count = r.count; return 0;
}
int64_t ref_convert_utf8_to_utf32(const char *impl, const char *in, size_t len, char32_t *dst) {
  auto *i = pick(impl); if (!i) return -1;
  return int64_t(i->convert_utf8_to_utf32(in, len, dst));
}
int64_t ref_count_utf16le(const char *impl, const char16_t *in, size_t len) {
  auto *i = pick(impl); if (!i) return -1;
  return int64_t(i->count_utf16le(in, len));
}
int64_t ref_utf8_length_from_utf16le(const char *impl, const char16_t *in, size_t len) {
  auto *i = pick(impl); if (!i) return -1;
  return int64_t(i->utf8_length_from_utf16le(in, len));
}
int ref_validate_utf16le_with_errors(const char *impl, const char16_t *in, size_t len, ref_result *out) {
  auto *i = pick(impl); if (!i) return -1;
  simdutf::result r = i->validate_utf16le_with_errors(in, len);
  out->error = int32_t(r.error); out->count = r.count; return 0;
}
int ref_convert_utf16le_to_utf8_with_errors(const char *impl, const char16_t *in, size_t len, char *dst, ref_result *out) {
  auto *i = pick(impl); if (!i) return -1;
  simdutf::result r = i->convert_utf16le_to_utf8_with_errors(in, len, dst);
  out->error = int32_t(r.error); out->count = r.count; return 0;
}
int64_t ref_convert_utf16le_to_utf8(const char *impl, const char16_t *in, size_t len, char *dst) {
  auto *i = pick(impl); if (!i) return -1;
  return int64_t(i->convert_utf16le_to_utf8(in, len, dst));
}
// UTF-16BE twins
int64_t ref_count_utf16be(const char *impl, const char16_t *in, size_t len) {
  auto *i = pick(impl); if (!i) return -1;
  return int64_t(i->count_utf16be(in, len));
}
int64_t ref_utf8_length_from_utf16be(const char *impl, const char16_t *in, size_t len) {
  auto *i = pick(impl); if (!i) return -1;
  return int64_t(i->utf8_length_from_utf16be(in, len));
}
int ref_validate_utf16be_with_errors(const char *impl, const char16_t *in, size_t len, ref_result *out) {
  auto *i = pick(impl); if (!i) return -1;
  simdutf::result r = i->validate_utf16be_with_errors(in, len);
  out->error = int32_t(r.error); out->count = r.count; return 0;
}
int ref_convert_utf16be_to_utf8_with_errors(const char *impl, const char16_t *in, size_t len, char *dst, ref_result *out) {
  auto *i = pick(impl); if (!i) return -1;
  simdutf::result r = i->convert_utf16be_to_utf8_with_errors(in, len, dst);
  out->error = int32_t(r.error); out->count = r.count; return 0;
}
int ref_convert_utf8_to_utf16be_with_errors(const char *impl, const char *in, size_t len, char16_t *dst, ref_result *out) {
  auto *i = pick(impl); if (!i) return -1;
  simdutf::result r = i->convert_utf8_to_utf16be_with_errors(in, len, dst);
  out->error = int32_t(r.error); out->count = r.count; return 0;
}
int ref_change_endianness_utf16(const char *impl, const char16_t *in, size_t len, char16_t *dst) {
  auto *i = pick(impl); if (!i) return -1;
  i->change_endianness_utf16(in, len, dst); return 0;
}
// UTF-32 family (second part)
int ref_validate_utf32_with_errors(const char *impl, const char32_t *in, size_t len, ref_result *out) {
  auto *i = pick(impl); if (!i) return -1;
  simdutf::result r = i->validate_utf32_with_errors(in, len);
  out->error = int32_t(r.error); out->count = r.count; return 0;
}
int64_t ref_utf8_length_from_utf32(const char *impl, const char32_t *in, size_t len) {
  auto *i = pick(impl); if (!i) return -1;
  return int64_t(i->utf8_length_from_utf32(in, len));
}
int64_t ref_utf16_length_from_utf32(const char *impl, const char32_t *in, size_t len) {
  auto *i = pick(impl); if (!i) return -1;
  return int64_t(i->utf16_length_from_utf32(in, len));
}
int ref_convert_utf32_to_utf8_with_errors(const char *impl, const char32_t *in, size_t len, char *dst, ref_result *out) {
  auto *i = pick(impl); if (!i) return -1;
  simdutf::result r = i->convert_utf32_to_utf8_with_errors(in, len, dst);
  out->error = int32_t(r.error); out->count = r.count; return 0;
}
int ref_convert_utf32_to_utf16_with_errors(const char *impl, int be, const char32_t *in, size_t len, char16_t *dst, ref_result *out) {
  auto *i = pick(impl); if (!i) return -1;
  simdutf::result r = be ? i->convert_utf32_to_utf16be_with_errors(in, len, dst) : i->convert_utf32_to_utf16le_with_errors(in, len, dst);
  out->error = int32_t(r.error); out->count = r.count; return 0;
}
int ref_convert_utf16_to_utf32_with_errors(const char *impl, int be, const char16_t *in, size_t len, char32_t *dst, ref_result *out) {
  auto *i = pick(impl); if (!i) return -1;
  simdutf::result r = be ? i->convert_utf16be_to_utf32_with_errors(in, len, dst) : i->convert_utf16le_to_utf32_with_errors(in, len, dst);
  out->error = int32_t(r.error); out->count = r.count; return 0;
}
// Latin-1 / ASCII family
int ref_validate_ascii_with_errors(const char *impl, const char *in, size_t len, ref_result *out) {
  auto *i = pick(impl); if (!i) return -1;
  simdutf::result r = i->validate_ascii_with_errors(in, len);
  out->error = int32_t(r.error); out->count = r.count; return 0;
}
int64_t ref_utf8_length_from_latin1(const char *impl, const char *in, size_t len) {
  auto *i = pick(impl); if (!i) return -1;
  return int64_t(i->utf8_length_from_latin1(in, len));
}
int64_t ref_latin1_length_from_utf8(const char *impl, const char *in, size_t len) {
  auto *i = pick(impl); if (!i) return -1;
  return int64_t(i->latin1_length_from_utf8(in, len));
}
int64_t ref_convert_latin1_to_utf8(const char *impl, const char *in, size_t len, char *dst) {
  auto *i = pick(impl); if (!i) return -1;
  return int64_t(i->convert_latin1_to_utf8(in, len, dst));
}
int64_t ref_convert_latin1_to_utf16(const char *impl, int be, const char *in, size_t len, char16_t *dst) {
  auto *i = pick(impl); if (!i) return -1;
  return int64_t(be ? i->convert_latin1_to_utf16be(in, len, dst) : i->convert_latin1_to_utf16le(in, len, dst));
}
int64_t ref_convert_latin1_to_utf32(const char *impl, const char *in, size_t len, char32_t *dst) {
  auto *i = pick(impl); if (!i) return -1;
  return int64_t(i->convert_latin1_to_utf32(in, len, dst));
}
int ref_convert_utf8_to_latin1_with_errors(const char *impl, const char *in, size_t len, char *dst, ref_result *out) {
  auto *i = pick(impl); if (!i) return -1;
  simdutf::result r = i->convert_utf8_to_latin1_with_errors(in, len, dst);
  out->error = int32_t(r.error); out->count = r.count; return 0;
}
int ref_convert_utf16_to_latin1_with_errors(const char *impl, int be, const char16_t *in, size_t len, char *dst, ref_result *out) {
  auto *i = pick(impl); if (!i) return -1;
  simdutf::result r = be ? i->convert_utf16be_to_latin1_with_errors(in, len, dst) : i->convert_utf16le_to_latin1_with_errors(in, len, dst);
  out->error = int32_t(r.error); out->count = r.count; return 0;
}
int ref_convert_utf32_to_latin1_with_errors(const char *impl, const char32_t *in, size_t len, char *dst, ref_result *out) {
  auto *i = pick(impl); if (!i) return -1;
  simdutf::result r = i->convert_utf32_to_latin1_with_errors(in, len, dst);
  out->error = int32_t(r.error); out->count = r.count; return 0;
}
// SURVEY.md §8f rank 4
int ref_to_well_formed_utf16(const char *impl, int be, const char16_t *in, size_t len, char16_t *dst) {
  auto *i = pick(impl); if (!i) return -1;
  if (be) i->to_well_formed_utf16be(in, len, dst); else i->to_well_formed_utf16le(in, len, dst);
  return 0;
}
int ref_detect_encodings(const char *impl, const char *in, size_t len) {
  auto *i = pick(impl); if (!i) return -1;
  return i->detect_encodings(in, len);
}
int64_t ref_maximal_binary_length_from_base64(const char *in, size_t len) {
  return int64_t(simdutf::maximal_binary_length_from_base64(in, len));
}
int ref_base64_to_binary(const char *impl, const char *in, size_t len, char *dst, uint64_t options, uint64_t last_chunk, ref_result *out) {
  auto *i = pick(impl); if (!i) return -1;
  simdutf::result r = i->base64_to_binary(in, len, dst, simdutf::base64_options(options),
                                          simdutf::last_chunk_handling_options(last_chunk));
  out->error = int32_t(r.error); out->count = r.count; return 0;
}
int ref_base64_to_binary_details(const char *impl, const char *in, size_t len, char *dst, uint64_t options, uint64_t last_chunk, ref_full_result *out) {
  auto *i = pick(impl); if (!i) return -1;
  simdutf::full_result r = i->base64_to_binary_details(in, len, dst, simdutf::base64_options(options),
                                          simdutf::last_chunk_handling_options(last_chunk));
  out->error = int32_t(r.error); out->input_count = r.input_count; out->output_count = r.output_count; return 0;
}
int64_t ref_base64_length_from_binary(size_t len, uint64_t options) {
  return int64_t(simdutf::base64_length_from_binary(len, simdutf::base64_options(options)));
}
int64_t ref_binary_to_base64(const char *impl, const char *in, size_t len, char *dst, uint64_t options) {
  auto *i = pick(impl); if (!i) return -1;
  return int64_t(i->binary_to_base64(in, len, dst, simdutf::base64_options(options)));
}
int64_t ref_trim_partial_utf8(const char *in, size_t len) {
  return int64_t(simdutf::trim_partial_utf8(in, len));
}

// ---------------------------------------------------------------------------
// Multi-threaded CPU baseline: the recipe of the reference's
// benchmarks/threaded.cpp:69-88 generalised to T threads — cut the input at
// code-point boundaries, size every chunk's output with
// utf16_length_from_utf8, then convert all chunks concurrently.  Used ONLY by
// bench.py's cpu_baseline / --impl reference legs.
// Returns total units written, or -1 on error / invalid input.
int64_t ref_mt_utf16_length_then_convert_utf8_to_utf16le(const char *impl, const char *in, size_t len,
                                                         char16_t *dst, int threads) {
  auto *i = pick(impl); if (!i) return -1;
  if (threads < 1) threads = 1;
  std::vector<size_t> cut(threads + 1, 0);
  cut[threads] = len;
  for (int t = 1; t < threads; t++) {
    size_t c = len / threads * t;
    while (c > 0 && (uint8_t(in[c]) & 0xC0) == 0x80) c--;  // back up to a lead (<=3 B on valid data)
    cut[t] = std::max(c, cut[t - 1]);
  }
  std::vector<size_t> units(threads, 0);
  {
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; t++)
      pool.emplace_back([&, t] { units[t] = i->utf16_length_from_utf8(in + cut[t], cut[t + 1] - cut[t]); });
    for (auto &th : pool) th.join();
  }
  std::vector<size_t> off(threads + 1, 0);
  for (int t = 0; t < threads; t++) off[t + 1] = off[t] + units[t];
  std::atomic<int> bad{0};
  {
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; t++)
      pool.emplace_back([&, t] {
        simdutf::result r = i->convert_utf8_to_utf16le_with_errors(in + cut[t], cut[t + 1] - cut[t], dst + off[t]);
        if (r.error) bad = 1;
      });
    for (auto &th : pool) th.join();
  }
  return bad ? -1 : int64_t(off[threads]);
}

// T-thread validate_utf8_with_errors over code-point-aligned chunks; first
// error = minimum over chunks (global position).
int ref_mt_validate_utf8_with_errors(const char *impl, const char *in, size_t len, int threads, ref_result *out) {
  auto *i = pick(impl); if (!i) return -1;
  if (threads < 1) threads = 1;
  std::vector<size_t> cut(threads + 1, 0);
  cut[threads] = len;
  for (int t = 1; t < threads; t++) {
    size_t c = len / threads * t;
    int back = 0;
    while (c > 0 && back < 3 && (uint8_t(in[c]) & 0xC0) == 0x80) { c--; back++; }
    cut[t] = std::max(c, cut[t - 1]);
  }
  std::vector<simdutf::result> res(threads);
  std::vector<std::thread> pool;
  for (int t = 0; t < threads; t++)
    pool.emplace_back([&, t] { res[t] = i->validate_utf8_with_errors(in + cut[t], cut[t + 1] - cut[t]); });
  for (auto &th : pool) th.join();
  out->error = 0; out->count = len;
  for (int t = 0; t < threads; t++) {
    if (res[t].error) { out->error = int32_t(res[t].error); out->count = cut[t] + res[t].count; break; }
  }
  return 0;
}

} // extern "C"
