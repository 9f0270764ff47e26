/* oracle/oracle.h — TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C restatement of the scalar semantics of the reference hot path
 * (WojciechMula/simdutf v7.0.0, /root/reference/src/scalar/ **).  It is the
 * CPU checker the CUDA kernels are compared with.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline/--impl reference legs may
 * load it; the product library (simdutf_b200/libsimdutf_b200.so) never links,
 * loads or calls anything in oracle/.
 *
 * Parity status: PINNED.  tests/test_oracle.py checks every function here
 * against (a) the golden vectors in tests/golden/ (known-answer tests
 * transcribed from the reference's own tests + outputs recorded from the
 * reference library itself by tests/golden/make_golden.py) and, when
 * oracle/_ref/libsimdutf_ref.so is present, (b) the unmodified reference
 * library's icelake / haswell / fallback kernels on seeded random inputs.
 */
#ifndef SIMDUTF_B200_ORACLE_H
#define SIMDUTF_B200_ORACLE_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* reference include/simdutf/error.h:5-32 */
enum oracle_error {
  ORACLE_SUCCESS = 0,
  ORACLE_HEADER_BITS = 1,
  ORACLE_TOO_SHORT = 2,
  ORACLE_TOO_LONG = 3,
  ORACLE_OVERLONG = 4,
  ORACLE_TOO_LARGE = 5,
  ORACLE_SURROGATE = 6,
  ORACLE_INVALID_BASE64_CHARACTER = 7,
  ORACLE_BASE64_INPUT_REMAINDER = 8,
  ORACLE_BASE64_EXTRA_BITS = 9,
  ORACLE_OUTPUT_BUFFER_TOO_SMALL = 10,
  ORACLE_OTHER = 11
};

/* reference include/simdutf/error.h:34-37 and :54-57 */
typedef struct { int32_t error; uint64_t count; } oracle_result;
typedef struct { int32_t error; uint64_t input_count; uint64_t output_count; } oracle_full_result;

/* base64_options / last_chunk_handling_options: reference
 * include/simdutf/implementation.h:2782-2811 */
enum { ORACLE_B64_DEFAULT = 0, ORACLE_B64_URL = 1, ORACLE_B64_REVERSE_PADDING = 2,
       ORACLE_B64_DEFAULT_ACCEPT_GARBAGE = 4, ORACLE_B64_URL_ACCEPT_GARBAGE = 5,
       ORACLE_B64_DEFAULT_OR_URL = 8, ORACLE_B64_DEFAULT_OR_URL_ACCEPT_GARBAGE = 12 };
enum { ORACLE_LOOSE = 0, ORACLE_STRICT = 1, ORACLE_STOP_BEFORE_PARTIAL = 2 };

oracle_result oracle_validate_utf8_with_errors(const uint8_t *in, size_t len);
int oracle_validate_utf8(const uint8_t *in, size_t len);
uint64_t oracle_count_utf8(const uint8_t *in, size_t len);
uint64_t oracle_utf16_length_from_utf8(const uint8_t *in, size_t len);
uint64_t oracle_utf32_length_from_utf8(const uint8_t *in, size_t len);
oracle_result oracle_convert_utf8_to_utf16le_with_errors(const uint8_t *in, size_t len, uint16_t *out);
uint64_t oracle_convert_utf8_to_utf16le(const uint8_t *in, size_t len, uint16_t *out);
oracle_result oracle_convert_utf8_to_utf32_with_errors(const uint8_t *in, size_t len, uint32_t *out);
uint64_t oracle_convert_utf8_to_utf32(const uint8_t *in, size_t len, uint32_t *out);

oracle_result oracle_validate_utf16le_with_errors(const uint16_t *in, size_t len);
uint64_t oracle_count_utf16le(const uint16_t *in, size_t len);
uint64_t oracle_utf8_length_from_utf16le(const uint16_t *in, size_t len);
uint64_t oracle_utf32_length_from_utf16le(const uint16_t *in, size_t len);
oracle_result oracle_convert_utf16le_to_utf8_with_errors(const uint16_t *in, size_t len, uint8_t *out);
uint64_t oracle_convert_utf16le_to_utf8(const uint16_t *in, size_t len, uint8_t *out);

/* UTF-16BE twins: same routines, units byte-swapped on the way in / out */
oracle_result oracle_validate_utf16be_with_errors(const uint16_t *in, size_t len);
uint64_t oracle_count_utf16be(const uint16_t *in, size_t len);
uint64_t oracle_utf8_length_from_utf16be(const uint16_t *in, size_t len);
uint64_t oracle_utf32_length_from_utf16be(const uint16_t *in, size_t len);
oracle_result oracle_convert_utf16be_to_utf8_with_errors(const uint16_t *in, size_t len, uint8_t *out);
oracle_result oracle_convert_utf8_to_utf16be_with_errors(const uint8_t *in, size_t len, uint16_t *out);
void oracle_change_endianness_utf16(const uint16_t *in, size_t len, uint16_t *out);

/* UTF-32 family (second part) */
oracle_result oracle_validate_utf32_with_errors(const uint32_t *in, size_t len);
uint64_t oracle_utf8_length_from_utf32(const uint32_t *in, size_t len);
uint64_t oracle_utf16_length_from_utf32(const uint32_t *in, size_t len);
oracle_result oracle_convert_utf32_to_utf8_with_errors(const uint32_t *in, size_t len, uint8_t *out);
oracle_result oracle_convert_utf32_to_utf16le_with_errors(const uint32_t *in, size_t len, uint16_t *out);
oracle_result oracle_convert_utf32_to_utf16be_with_errors(const uint32_t *in, size_t len, uint16_t *out);
oracle_result oracle_convert_utf16le_to_utf32_with_errors(const uint16_t *in, size_t len, uint32_t *out);
oracle_result oracle_convert_utf16be_to_utf32_with_errors(const uint16_t *in, size_t len, uint32_t *out);

/* Latin-1 / ASCII family */
oracle_result oracle_validate_ascii_with_errors(const uint8_t *in, size_t len);
uint64_t oracle_utf8_length_from_latin1(const uint8_t *in, size_t len);
uint64_t oracle_convert_latin1_to_utf8(const uint8_t *in, size_t len, uint8_t *out);
uint64_t oracle_convert_latin1_to_utf16(const uint8_t *in, size_t len, uint16_t *out, int be);
uint64_t oracle_convert_latin1_to_utf32(const uint8_t *in, size_t len, uint32_t *out);
oracle_result oracle_convert_utf8_to_latin1_with_errors(const uint8_t *in, size_t len, uint8_t *out);
oracle_result oracle_convert_utf16_to_latin1_with_errors(const uint16_t *in, size_t len, uint8_t *out, int be);
oracle_result oracle_convert_utf32_to_latin1_with_errors(const uint32_t *in, size_t len, uint8_t *out);

/* SURVEY.md §8f rank 4 */
void oracle_to_well_formed_utf16(const uint16_t *in, size_t len, uint16_t *out, int be);
int oracle_detect_encodings(const uint8_t *in, size_t len);

uint64_t oracle_maximal_binary_length_from_base64(const uint8_t *in, size_t len);
oracle_full_result oracle_base64_to_binary_details(const uint8_t *in, size_t len, uint8_t *out,
                                                   uint64_t options, uint64_t last_chunk);
oracle_result oracle_base64_to_binary(const uint8_t *in, size_t len, uint8_t *out,
                                      uint64_t options, uint64_t last_chunk);
uint64_t oracle_base64_length_from_binary(size_t len, uint64_t options);
uint64_t oracle_binary_to_base64(const uint8_t *in, size_t len, uint8_t *out, uint64_t options);

uint64_t oracle_trim_partial_utf8(const uint8_t *in, size_t len);
uint64_t oracle_trim_partial_utf16le(const uint16_t *in, size_t len);

#ifdef __cplusplus
}
#endif
#endif
