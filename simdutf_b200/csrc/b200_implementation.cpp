// b200_implementation.cpp — bodies of simdutf::b200::implementation (see b200_implementation.h).
//
// Each hot-path virtual is one call into the C ABI of libsimdutf_b200.so with the caller's HOST pointers;
// staging to the device, the kernels and the copy back all happen behind b200_host_*.  Error conventions are
// the reference's: `_with_errors` return result{error, position | count}; the plain converters return the count
// or 0 on any error (reference include/simdutf/implementation.h:3705-3706); an infrastructure failure (no
// device, CUDA error) becomes error_code::OTHER / 0 / false, never an exception, a message or an abort
// (reference include/simdutf/error.h:31, src/implementation.cpp:819-822, CMakeLists.txt:173-214).
#include "b200_implementation.h"

#include "simdutf_b200.h"

#ifdef SIMDUTF_INTERNAL_TESTS
  #include <cstdio>
  #include <cstdlib>
  #include <vector>
#endif

namespace simdutf {
namespace b200 {

namespace {
simdutf_really_inline result to_result(int status, const b200_result &r) {
  if (status != 0) return result(error_code::OTHER, 0);
  return result(error_code(r.error), size_t(r.count));
}
simdutf_really_inline const uint16_t *u16(const char16_t *p) { return reinterpret_cast<const uint16_t *>(p); }
} // namespace

uint32_t implementation::required_instruction_sets() const {
  return b200_device_count() > 0 ? 0u : 0x80000000u;
}

// ---- UTF-8 validation (reference include/simdutf/implementation.h:3378-3396) ----
bool implementation::validate_utf8(const char *buf, size_t len) const noexcept {
  b200_result r;
  return b200_host_validate_utf8_with_errors(buf, len, &r) == 0 && r.error == B200_SUCCESS;
}
result implementation::validate_utf8_with_errors(const char *buf, size_t len) const noexcept {
  b200_result r;
  return to_result(b200_host_validate_utf8_with_errors(buf, len, &r), r);
}

// ---- UTF-8 counting (:3863, :3882, :4802) ----
size_t implementation::count_utf8(const char *input, size_t length) const noexcept {
  uint64_t n = 0;
  return b200_host_count_utf8(input, length, &n) == 0 ? size_t(n) : 0;
}
size_t implementation::utf32_length_from_utf8(const char *input, size_t length) const noexcept {
  return count_utf8(input, length);
}
size_t implementation::utf16_length_from_utf8(const char *input, size_t length) const noexcept {
  uint64_t n = 0;
  return b200_host_utf16_length_from_utf8(input, length, &n) == 0 ? size_t(n) : 0;
}

// ---- UTF-8 -> UTF-16LE (:3709, :3743-3745, :3815) ----
result implementation::convert_utf8_to_utf16le_with_errors(const char *input, size_t length,
                                                           char16_t *utf16_output) const noexcept {
  b200_result r;
  return to_result(b200_host_convert_utf8_to_utf16le(input, length, reinterpret_cast<uint16_t *>(utf16_output), &r), r);
}
size_t implementation::convert_utf8_to_utf16le(const char *input, size_t length, char16_t *utf16_output) const noexcept {
  const result r = convert_utf8_to_utf16le_with_errors(input, length, utf16_output);
  return r.error ? 0 : r.count;
}
size_t implementation::convert_valid_utf8_to_utf16le(const char *input, size_t length,
                                                     char16_t *utf16_output) const noexcept {
  return convert_utf8_to_utf16le(input, length, utf16_output);
}

// ---- UTF-8 -> UTF-32 (:3781, :3799-3800) ----
result implementation::convert_utf8_to_utf32_with_errors(const char *input, size_t length,
                                                         char32_t *utf32_output) const noexcept {
  b200_result r;
  return to_result(b200_host_convert_utf8_to_utf32(input, length, reinterpret_cast<uint32_t *>(utf32_output), &r), r);
}
size_t implementation::convert_utf8_to_utf32(const char *input, size_t length, char32_t *utf32_output) const noexcept {
  const result r = convert_utf8_to_utf32_with_errors(input, length, utf32_output);
  return r.error ? 0 : r.count;
}
size_t implementation::convert_valid_utf8_to_utf32(const char *input, size_t length,
                                                   char32_t *utf32_output) const noexcept {
  return convert_utf8_to_utf32(input, length, utf32_output);
}

// ---- UTF-16LE counting / validation (:3465-3483, :4277-4278, :4767) ----
size_t implementation::count_utf16le(const char16_t *input, size_t length) const noexcept {
  uint64_t n = 0;
  return b200_host_count_utf16le(u16(input), length, &n) == 0 ? size_t(n) : 0;
}
size_t implementation::utf32_length_from_utf16le(const char16_t *input, size_t length) const noexcept {
  return count_utf16le(input, length);
}
size_t implementation::utf8_length_from_utf16le(const char16_t *input, size_t length) const noexcept {
  uint64_t n = 0;
  return b200_host_utf8_length_from_utf16le(u16(input), length, &n) == 0 ? size_t(n) : 0;
}
bool implementation::validate_utf16le(const char16_t *buf, size_t len) const noexcept {
  b200_result r;
  return b200_host_validate_utf16le_with_errors(u16(buf), len, &r) == 0 && r.error == B200_SUCCESS;
}
result implementation::validate_utf16le_with_errors(const char16_t *buf, size_t len) const noexcept {
  b200_result r;
  return to_result(b200_host_validate_utf16le_with_errors(u16(buf), len, &r), r);
}

// ---- UTF-16LE -> UTF-8 (:4038, :4079-4080) ----
result implementation::convert_utf16le_to_utf8_with_errors(const char16_t *input, size_t length,
                                                           char *utf8_buffer) const noexcept {
  b200_result r;
  return to_result(b200_host_convert_utf16le_to_utf8(u16(input), length, utf8_buffer, &r), r);
}
size_t implementation::convert_utf16le_to_utf8(const char16_t *input, size_t length, char *utf8_buffer) const noexcept {
  const result r = convert_utf16le_to_utf8_with_errors(input, length, utf8_buffer);
  return r.error ? 0 : r.count;
}
size_t implementation::convert_valid_utf16le_to_utf8(const char16_t *input, size_t length,
                                                     char *utf8_buffer) const noexcept {
  return convert_utf16le_to_utf8(input, length, utf8_buffer);
}

// ---- UTF-16BE twins (reference include/simdutf/implementation.h:3443-3569, 3727-3760,
//      4058-4101, 4299-4368, 4567-4584, 4783-4800): the same kernels with big-endian units ----
size_t implementation::count_utf16be(const char16_t *input, size_t length) const noexcept {
  uint64_t n = 0;
  return b200_host_count_utf16be(u16(input), length, &n) == 0 ? size_t(n) : 0;
}
size_t implementation::utf32_length_from_utf16be(const char16_t *input, size_t length) const noexcept {
  return count_utf16be(input, length);
}
size_t implementation::utf8_length_from_utf16be(const char16_t *input, size_t length) const noexcept {
  uint64_t n = 0;
  return b200_host_utf8_length_from_utf16be(u16(input), length, &n) == 0 ? size_t(n) : 0;
}
bool implementation::validate_utf16be(const char16_t *buf, size_t len) const noexcept {
  b200_result r;
  return b200_host_validate_utf16be_with_errors(u16(buf), len, &r) == 0 && r.error == B200_SUCCESS;
}
result implementation::validate_utf16be_with_errors(const char16_t *buf, size_t len) const noexcept {
  b200_result r;
  return to_result(b200_host_validate_utf16be_with_errors(u16(buf), len, &r), r);
}
result implementation::convert_utf8_to_utf16be_with_errors(const char *input, size_t length,
                                                           char16_t *utf16_output) const noexcept {
  b200_result r;
  return to_result(b200_host_convert_utf8_to_utf16be(input, length, reinterpret_cast<uint16_t *>(utf16_output), &r), r);
}
size_t implementation::convert_utf8_to_utf16be(const char *input, size_t length, char16_t *utf16_output) const noexcept {
  const result r = convert_utf8_to_utf16be_with_errors(input, length, utf16_output);
  return r.error ? 0 : r.count;
}
size_t implementation::convert_valid_utf8_to_utf16be(const char *input, size_t length,
                                                     char16_t *utf16_buffer) const noexcept {
  return convert_utf8_to_utf16be(input, length, utf16_buffer);
}
result implementation::convert_utf16be_to_utf8_with_errors(const char16_t *input, size_t length,
                                                           char *utf8_buffer) const noexcept {
  b200_result r;
  return to_result(b200_host_convert_utf16be_to_utf8(u16(input), length, utf8_buffer, &r), r);
}
size_t implementation::convert_utf16be_to_utf8(const char16_t *input, size_t length, char *utf8_buffer) const noexcept {
  const result r = convert_utf16be_to_utf8_with_errors(input, length, utf8_buffer);
  return r.error ? 0 : r.count;
}
size_t implementation::convert_valid_utf16be_to_utf8(const char16_t *input, size_t length,
                                                     char *utf8_buffer) const noexcept {
  return convert_utf16be_to_utf8(input, length, utf8_buffer);
}
void implementation::change_endianness_utf16(const char16_t *input, size_t length, char16_t *output) const noexcept {
  b200_result r;
  (void)b200_host_change_endianness_utf16(u16(input), length, reinterpret_cast<uint16_t *>(output), &r);
}

// ---- base64 decode, char input (:4866-4870, :4902-4906); result derived as in include/simdutf/error.h:66-73 ----
full_result implementation::base64_to_binary_details(const char *input, size_t length, char *output,
                                                     base64_options options,
                                                     last_chunk_handling_options last_chunk_options) const noexcept {
  b200_full_result r;
  if (b200_host_base64_to_binary(input, length, output, uint64_t(options), uint64_t(last_chunk_options), &r) != 0) {
    return full_result(error_code::OTHER, 0, 0);
  }
  return full_result(error_code(r.error), size_t(r.input_count), size_t(r.output_count));
}
result implementation::base64_to_binary(const char *input, size_t length, char *output, base64_options options,
                                        last_chunk_handling_options last_chunk_options) const noexcept {
  return base64_to_binary_details(input, length, output, options, last_chunk_options);
}

// ---- base64 decode, char16_t input (:4922-4939, :4976-5014) ----
full_result implementation::base64_to_binary_details(const char16_t *input, size_t length, char *output,
                                                     base64_options options,
                                                     last_chunk_handling_options last_chunk_options) const noexcept {
  b200_full_result r;
  if (b200_host_base64_to_binary_utf16(u16(input), length, output, uint64_t(options), uint64_t(last_chunk_options), &r) != 0) {
    return full_result(error_code::OTHER, 0, 0);
  }
  return full_result(error_code(r.error), size_t(r.input_count), size_t(r.output_count));
}
result implementation::base64_to_binary(const char16_t *input, size_t length, char *output, base64_options options,
                                        last_chunk_handling_options last_chunk_options) const noexcept {
  return base64_to_binary_details(input, length, output, options, last_chunk_options);
}

// ---- binary_to_base64 (:4941-4960) ----
size_t implementation::binary_to_base64(const char *input, size_t length, char *output,
                                        base64_options options) const noexcept {
  b200_result r;
  if (b200_host_binary_to_base64(input, length, output, uint64_t(options) & 3u, &r) != 0) return 0;
  return size_t(r.count);
}

// ---- UTF-32 family (second part; reference include/simdutf/implementation.h:3533-3569,
//      4103-4259, 4370-4565, 4586-4620) ----
static inline const uint32_t *u32(const char32_t *p) { return reinterpret_cast<const uint32_t *>(p); }
result implementation::validate_utf32_with_errors(const char32_t *buf, size_t len) const noexcept {
  b200_result r;
  return to_result(b200_host_validate_utf32_with_errors(u32(buf), len, &r), r);
}
bool implementation::validate_utf32(const char32_t *buf, size_t len) const noexcept {
  return validate_utf32_with_errors(buf, len).error == error_code::SUCCESS;
}
size_t implementation::utf8_length_from_utf32(const char32_t *input, size_t length) const noexcept {
  uint64_t n = 0;
  return b200_host_utf8_length_from_utf32(u32(input), length, &n) == 0 ? size_t(n) : 0;
}
size_t implementation::utf16_length_from_utf32(const char32_t *input, size_t length) const noexcept {
  uint64_t n = 0;
  return b200_host_utf16_length_from_utf32(u32(input), length, &n) == 0 ? size_t(n) : 0;
}
result implementation::convert_utf32_to_utf8_with_errors(const char32_t *input, size_t length,
                                                         char *utf8_buffer) const noexcept {
  b200_result r;
  return to_result(b200_host_convert_utf32_to_utf8(u32(input), length, utf8_buffer, &r), r);
}
size_t implementation::convert_utf32_to_utf8(const char32_t *input, size_t length, char *utf8_buffer) const noexcept {
  const result r = convert_utf32_to_utf8_with_errors(input, length, utf8_buffer);
  return r.error ? 0 : r.count;
}
size_t implementation::convert_valid_utf32_to_utf8(const char32_t *input, size_t length,
                                                   char *utf8_buffer) const noexcept {
  return convert_utf32_to_utf8(input, length, utf8_buffer);
}
result implementation::convert_utf32_to_utf16le_with_errors(const char32_t *input, size_t length,
                                                            char16_t *utf16_buffer) const noexcept {
  b200_result r;
  return to_result(b200_host_convert_utf32_to_utf16le(u32(input), length, reinterpret_cast<uint16_t *>(utf16_buffer), &r), r);
}
size_t implementation::convert_utf32_to_utf16le(const char32_t *input, size_t length,
                                                char16_t *utf16_buffer) const noexcept {
  const result r = convert_utf32_to_utf16le_with_errors(input, length, utf16_buffer);
  return r.error ? 0 : r.count;
}
size_t implementation::convert_valid_utf32_to_utf16le(const char32_t *input, size_t length,
                                                      char16_t *utf16_buffer) const noexcept {
  return convert_utf32_to_utf16le(input, length, utf16_buffer);
}
result implementation::convert_utf32_to_utf16be_with_errors(const char32_t *input, size_t length,
                                                            char16_t *utf16_buffer) const noexcept {
  b200_result r;
  return to_result(b200_host_convert_utf32_to_utf16be(u32(input), length, reinterpret_cast<uint16_t *>(utf16_buffer), &r), r);
}
size_t implementation::convert_utf32_to_utf16be(const char32_t *input, size_t length,
                                                char16_t *utf16_buffer) const noexcept {
  const result r = convert_utf32_to_utf16be_with_errors(input, length, utf16_buffer);
  return r.error ? 0 : r.count;
}
size_t implementation::convert_valid_utf32_to_utf16be(const char32_t *input, size_t length,
                                                      char16_t *utf16_buffer) const noexcept {
  return convert_utf32_to_utf16be(input, length, utf16_buffer);
}
result implementation::convert_utf16le_to_utf32_with_errors(const char16_t *input, size_t length,
                                                            char32_t *utf32_buffer) const noexcept {
  b200_result r;
  return to_result(b200_host_convert_utf16le_to_utf32(u16(input), length, reinterpret_cast<uint32_t *>(utf32_buffer), &r), r);
}
size_t implementation::convert_utf16le_to_utf32(const char16_t *input, size_t length,
                                                char32_t *utf32_buffer) const noexcept {
  const result r = convert_utf16le_to_utf32_with_errors(input, length, utf32_buffer);
  return r.error ? 0 : r.count;
}
size_t implementation::convert_valid_utf16le_to_utf32(const char16_t *input, size_t length,
                                                      char32_t *utf32_buffer) const noexcept {
  return convert_utf16le_to_utf32(input, length, utf32_buffer);
}
result implementation::convert_utf16be_to_utf32_with_errors(const char16_t *input, size_t length,
                                                            char32_t *utf32_buffer) const noexcept {
  b200_result r;
  return to_result(b200_host_convert_utf16be_to_utf32(u16(input), length, reinterpret_cast<uint32_t *>(utf32_buffer), &r), r);
}
size_t implementation::convert_utf16be_to_utf32(const char16_t *input, size_t length,
                                                char32_t *utf32_buffer) const noexcept {
  const result r = convert_utf16be_to_utf32_with_errors(input, length, utf32_buffer);
  return r.error ? 0 : r.count;
}
size_t implementation::convert_valid_utf16be_to_utf32(const char16_t *input, size_t length,
                                                      char32_t *utf32_buffer) const noexcept {
  return convert_utf16be_to_utf32(input, length, utf32_buffer);
}

// ---- Latin-1 / ASCII family (reference include/simdutf/implementation.h:3409-3425, 3583-3692,
//      3885-4019, 4299-4368, 4586-4649) ----
result implementation::validate_ascii_with_errors(const char *buf, size_t len) const noexcept {
  b200_result r;
  return to_result(b200_host_validate_ascii_with_errors(buf, len, &r), r);
}
bool implementation::validate_ascii(const char *buf, size_t len) const noexcept {
  return validate_ascii_with_errors(buf, len).error == error_code::SUCCESS;
}
size_t implementation::utf8_length_from_latin1(const char *input, size_t length) const noexcept {
  uint64_t n = 0;
  return b200_host_utf8_length_from_latin1(input, length, &n) == 0 ? size_t(n) : 0;
}
size_t implementation::latin1_length_from_utf8(const char *input, size_t length) const noexcept {
  return count_utf8(input, length);
}
size_t implementation::convert_latin1_to_utf8(const char *input, size_t length, char *utf8_output) const noexcept {
  b200_result r;
  return b200_host_convert_latin1_to_utf8(input, length, utf8_output, &r) == 0 ? size_t(r.count) : 0;
}
size_t implementation::convert_latin1_to_utf16le(const char *input, size_t length, char16_t *utf16_output) const noexcept {
  b200_result r;
  return b200_host_convert_latin1_to_utf16le(input, length, reinterpret_cast<uint16_t *>(utf16_output), &r) == 0 ? size_t(r.count) : 0;
}
size_t implementation::convert_latin1_to_utf16be(const char *input, size_t length, char16_t *utf16_output) const noexcept {
  b200_result r;
  return b200_host_convert_latin1_to_utf16be(input, length, reinterpret_cast<uint16_t *>(utf16_output), &r) == 0 ? size_t(r.count) : 0;
}
size_t implementation::convert_latin1_to_utf32(const char *input, size_t length, char32_t *utf32_buffer) const noexcept {
  b200_result r;
  return b200_host_convert_latin1_to_utf32(input, length, reinterpret_cast<uint32_t *>(utf32_buffer), &r) == 0 ? size_t(r.count) : 0;
}
result implementation::convert_utf8_to_latin1_with_errors(const char *input, size_t length, char *output) const noexcept {
  b200_result r;
  return to_result(b200_host_convert_utf8_to_latin1(input, length, output, &r), r);
}
size_t implementation::convert_utf8_to_latin1(const char *input, size_t length, char *output) const noexcept {
  const result r = convert_utf8_to_latin1_with_errors(input, length, output);
  return r.error ? 0 : r.count;
}
size_t implementation::convert_valid_utf8_to_latin1(const char *input, size_t length, char *output) const noexcept {
  return convert_utf8_to_latin1(input, length, output);
}
result implementation::convert_utf16le_to_latin1_with_errors(const char16_t *input, size_t length, char *output) const noexcept {
  b200_result r;
  return to_result(b200_host_convert_utf16le_to_latin1(u16(input), length, output, &r), r);
}
size_t implementation::convert_utf16le_to_latin1(const char16_t *input, size_t length, char *output) const noexcept {
  const result r = convert_utf16le_to_latin1_with_errors(input, length, output);
  return r.error ? 0 : r.count;
}
size_t implementation::convert_valid_utf16le_to_latin1(const char16_t *input, size_t length, char *output) const noexcept {
  return convert_utf16le_to_latin1(input, length, output);
}
result implementation::convert_utf16be_to_latin1_with_errors(const char16_t *input, size_t length, char *output) const noexcept {
  b200_result r;
  return to_result(b200_host_convert_utf16be_to_latin1(u16(input), length, output, &r), r);
}
size_t implementation::convert_utf16be_to_latin1(const char16_t *input, size_t length, char *output) const noexcept {
  const result r = convert_utf16be_to_latin1_with_errors(input, length, output);
  return r.error ? 0 : r.count;
}
size_t implementation::convert_valid_utf16be_to_latin1(const char16_t *input, size_t length, char *output) const noexcept {
  return convert_utf16be_to_latin1(input, length, output);
}
result implementation::convert_utf32_to_latin1_with_errors(const char32_t *input, size_t length, char *output) const noexcept {
  b200_result r;
  return to_result(b200_host_convert_utf32_to_latin1(u32(input), length, output, &r), r);
}
size_t implementation::convert_utf32_to_latin1(const char32_t *input, size_t length, char *output) const noexcept {
  const result r = convert_utf32_to_latin1_with_errors(input, length, output);
  return r.error ? 0 : r.count;
}
size_t implementation::convert_valid_utf32_to_latin1(const char32_t *input, size_t length, char *output) const noexcept {
  return convert_utf32_to_latin1(input, length, output);
}

// ---- SURVEY.md §8f rank 4 (reference include/simdutf/implementation.h:3344-3354, 3498-3531) ----
void implementation::to_well_formed_utf16le(const char16_t *input, size_t len, char16_t *output) const noexcept {
  b200_result r;
  (void)b200_host_to_well_formed_utf16le(u16(input), len, reinterpret_cast<uint16_t *>(output), &r);
}
void implementation::to_well_formed_utf16be(const char16_t *input, size_t len, char16_t *output) const noexcept {
  b200_result r;
  (void)b200_host_to_well_formed_utf16be(u16(input), len, reinterpret_cast<uint16_t *>(output), &r);
}
int implementation::detect_encodings(const char *input, size_t length) const noexcept {
  uint64_t bits = 0;
  return b200_host_detect_encodings(input, length, &bits) == 0 ? int(bits) : 0;
}

#ifdef SIMDUTF_INTERNAL_TESTS
// ---- internal_tests() (reference include/simdutf/implementation.h:5036; src/ppc64/implementation.cpp:898-913 is the
// reference's own example: procedures abort() on failure).  Test-only code behind the reference's developer flag. ----
namespace {
void check(bool ok, const char *what) {
  if (!ok) {
    fprintf(stderr, "b200 internal test failed: %s (%s)\n", what, b200_last_error());
    abort();
  }
}
std::vector<char> mixed_text(size_t n) {  // valid UTF-8, 1-4-byte characters in rotation
  static const char *const chars[4] = {"a", "\xC3\xA9", "\xE4\xB8\xAD", "\xF0\x9F\x98\x80"};
  std::vector<char> v;
  v.reserve(n + 4);
  for (size_t i = 0; v.size() + 4 <= n; i++) {
    const char *c = chars[(i * 7 + i / 5) & 3];
    while (*c) v.push_back(*c++);
  }
  return v;
}
void host_path_over_all_devices(const simdutf::implementation &impl) {
  const int ndev = b200_device_count();
  const std::vector<char> text = mixed_text(size_t(96) << 20);
  const size_t units = impl.utf16_length_from_utf8(text.data(), text.size());
  std::vector<char16_t> one(units + 8, 0x5A5A), all(units + 8, 0x5A5A);
  check(b200_host_set_devices(1) == 0, "set_devices(1)");
  const result r1 = impl.convert_utf8_to_utf16le_with_errors(text.data(), text.size(), one.data());
  check(b200_host_set_devices(ndev) == 0, "set_devices(all)");
  const result rn = impl.convert_utf8_to_utf16le_with_errors(text.data(), text.size(), all.data());
  check(impl.utf16_length_from_utf8(text.data(), text.size()) == units, "length over all devices");
  b200_host_set_devices(1);
  check(r1.error == SUCCESS && r1.count == units && rn.error == SUCCESS && rn.count == units, "result over all devices");
  check(one == all && all[units] == 0x5A5A, "output over all devices equals output of one device");
}
void shard_helpers_and_combine(const simdutf::implementation &impl) {
  std::vector<char> text = mixed_text(size_t(3) << 20);
  text[(2u << 20) + 12345] = char(0xFF);
  const result whole = impl.validate_utf8_with_errors(text.data(), text.size());
  // cut in three at k*N/3 backed up with the trim helper, validate shard by shard, fold: first error in buffer order
  size_t cuts[4] = {0, 0, 0, text.size()};
  for (int k = 1; k < 3; k++) cuts[k] = b200_host_trim_partial_utf8(text.data(), text.size() * k / 3);
  result folded(SUCCESS, text.size());
  for (int k = 2; k >= 0; k--) {
    const result r = impl.validate_utf8_with_errors(text.data() + cuts[k], cuts[k + 1] - cuts[k]);
    if (r.error) folded = result(r.error, cuts[k] + r.count);
  }
  check(folded.error == whole.error && folded.count == whole.count && whole.error == HEADER_BITS, "sharded validate equals whole-buffer validate");
}
} // namespace
std::vector<simdutf::implementation::TestProcedure> implementation::internal_tests() const {
  return {TestProcedure{"b200_host_path_over_all_devices", host_path_over_all_devices},
          TestProcedure{"b200_shard_helpers_and_combine", shard_helpers_and_combine}};
}
#endif // SIMDUTF_INTERNAL_TESTS

// ---- anything the reference adds later: its "unsupported" answers (generated; empty against this reference) ----
#include "b200_stubs.inc"

} // namespace b200
} // namespace simdutf
