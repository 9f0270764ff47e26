/* simdutf_b200.h — C ABI of the B200 (sm_100a) backend for simdutf's hot path.
 *
 * This is the drop-in boundary: plain pointers and sizes, no C++/torch types.
 * Every entry point names the virtual of `class simdutf::implementation`
 * (reference include/simdutf/implementation.h) it stands behind; the C++
 * subclass `simdutf::b200::implementation` (simdutf_b200/csrc/b200_implementation.cpp)
 * forwards each of those virtuals to the matching `b200_host_*` function and
 * nothing else.  See INTEGRATION.md for the reference-side binding.
 *
 * Three flavours per operation:
 *   b200_<op>_async(d_in, ..., d_result, stream)
 *       DEVICE pointers in and out, result written to a DEVICE slot when the
 *       stream reaches it; no host synchronisation, graph-capturable.
 *   b200_<op>(d_in, ..., h_result, stream)
 *       DEVICE data pointers, result returned to the HOST (synchronises stream).
 *   b200_host_<op>(h_in, ..., h_result)
 *       HOST pointers (pinned or pageable): staged through the calling thread's
 *       device buffers in pipelined segments, over one or several devices
 *       (b200_host_set_devices); inputs of a few KiB take a zero-copy path.
 *       DEVICE pointers are recognised (cudaPointerGetAttributes) and processed
 *       in place on their device.  This is what the C++ virtuals call.
 *
 * Return value of every function: 0 on success, otherwise a CUDA error code
 * (cudaError_t, > 0) or a negative B200_E_* code.  simdutf-level outcomes
 * (error_code + position/count) travel in the result structs exactly as in the
 * reference (include/simdutf/error.h:5-74).  The library never prints, aborts
 * or throws (reference CMakeLists.txt:173-214), and it has NO CPU fallback: with
 * no sm_100 device present every compute entry point returns B200_E_NO_DEVICE.
 *
 * Lengths are in CODE UNITS of the input encoding (bytes for UTF-8 / Latin-1 /
 * base64, 16-bit units for UTF-16, 32-bit units for UTF-32), as in the reference.
 *
 * Coverage: the hot path of SURVEY.md §8a and all of §8f (UTF-16BE twins, UTF-32
 * family, Latin-1 / ASCII, base64 encode and char16_t decode, to_well_formed_utf16,
 * detect_encodings, change_endianness_utf16): every pure virtual of
 * simdutf::implementation has an entry point here.
 */
#ifndef SIMDUTF_B200_H
#define SIMDUTF_B200_H

#include <stddef.h>
#include <stdint.h>

#if defined(__GNUC__)
#define SIMDUTF_B200_API __attribute__((visibility("default")))
#else
#define SIMDUTF_B200_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

#define B200_E_NO_DEVICE (-1) /* no usable sm_100 device / CUDA not initialisable */
#define B200_E_BAD_ARGUMENT (-2)
#define B200_E_NO_MEMORY (-3)

/* simdutf::error_code, reference include/simdutf/error.h:5-32 (same values). */
enum b200_error_code {
  B200_SUCCESS = 0,
  B200_HEADER_BITS = 1,
  B200_TOO_SHORT = 2,
  B200_TOO_LONG = 3,
  B200_OVERLONG = 4,
  B200_TOO_LARGE = 5,
  B200_SURROGATE = 6,
  B200_INVALID_BASE64_CHARACTER = 7,
  B200_BASE64_INPUT_REMAINDER = 8,
  B200_BASE64_EXTRA_BITS = 9,
  B200_OUTPUT_BUFFER_TOO_SMALL = 10,
  B200_OTHER = 11
};

/* simdutf::result, reference include/simdutf/error.h:34-37 ({enum; size_t} = 16 B on LP64). */
typedef struct b200_result {
  int32_t error;
  uint32_t reserved_;
  uint64_t count;
} b200_result;

/* simdutf::full_result, reference include/simdutf/error.h:54-57 (24 B on LP64). */
typedef struct b200_full_result {
  int32_t error;
  uint32_t reserved_;
  uint64_t input_count;
  uint64_t output_count;
} b200_full_result;

/* simdutf::base64_options / last_chunk_handling_options,
 * reference include/simdutf/implementation.h:2782-2811 (same values). */
enum {
  B200_BASE64_DEFAULT = 0,
  B200_BASE64_URL = 1,
  B200_BASE64_REVERSE_PADDING = 2,
  B200_BASE64_DEFAULT_ACCEPT_GARBAGE = 4,
  B200_BASE64_URL_ACCEPT_GARBAGE = 5,
  B200_BASE64_DEFAULT_OR_URL = 8,
  B200_BASE64_DEFAULT_OR_URL_ACCEPT_GARBAGE = 12
};
enum { B200_LOOSE = 0, B200_STRICT = 1, B200_STOP_BEFORE_PARTIAL = 2 };

/* ------------------------------------------------------------------------- */
/* Runtime                                                                    */
/* ------------------------------------------------------------------------- */

/* Number of usable compute-capability-10.x devices (0 if none / no driver).
 * Backs implementation::required_instruction_sets() / supported_by_runtime_system()
 * (reference include/simdutf/implementation.h:3344-3366, src/implementation.cpp:35-41). */
SIMDUTF_B200_API int b200_device_count(void);
/* Select the device used by the calling thread's subsequent b200_* calls (default 0). */
SIMDUTF_B200_API int b200_set_device(int device);
SIMDUTF_B200_API int b200_get_device(void);
/* "b200" / description string, for implementation::name()/description(). */
SIMDUTF_B200_API const char *b200_name(void);
SIMDUTF_B200_API const char *b200_description(void);
/* Number of kernel launches issued by this library since load (all threads); bench.py's gpu_launches. */
SIMDUTF_B200_API uint64_t b200_launch_count(void);
/* Text of the last failure on this thread ("" if none). Never printed by the library. */
SIMDUTF_B200_API const char *b200_last_error(void);
/* Pinned host memory helpers for callers that want the fast host path. */
SIMDUTF_B200_API int b200_host_alloc(void **ptr, size_t bytes);
SIMDUTF_B200_API int b200_host_free(void *ptr);

/* ------------------------------------------------------------------------- */
/* UTF-8 validation — implementation::validate_utf8_with_errors               */
/* (reference include/simdutf/implementation.h:3395-3396; semantics            */
/*  src/scalar/utf8.h:102-200) and ::validate_utf8 (:3378-3379).               */
/* result = {SUCCESS, len} or {first error, byte index}.                       */
/* ------------------------------------------------------------------------- */
SIMDUTF_B200_API int b200_validate_utf8_with_errors_async(const char *d_in, size_t len, b200_result *d_res, void *stream);
SIMDUTF_B200_API int b200_validate_utf8_with_errors(const char *d_in, size_t len, b200_result *h_res, void *stream);
SIMDUTF_B200_API int b200_host_validate_utf8_with_errors(const char *h_in, size_t len, b200_result *h_res);

/* ------------------------------------------------------------------------- */
/* UTF-8 counting — implementation::count_utf8 (:4802) == utf32_length_from_utf8 */
/* (:3882); implementation::utf16_length_from_utf8 (:3863).                    */
/* Semantics src/scalar/utf8.h:230-255; never validate.                        */
/* ------------------------------------------------------------------------- */
SIMDUTF_B200_API int b200_count_utf8_async(const char *d_in, size_t len, uint64_t *d_count, void *stream);
SIMDUTF_B200_API int b200_count_utf8(const char *d_in, size_t len, uint64_t *h_count, void *stream);
SIMDUTF_B200_API int b200_host_count_utf8(const char *h_in, size_t len, uint64_t *h_count);
SIMDUTF_B200_API int b200_utf16_length_from_utf8_async(const char *d_in, size_t len, uint64_t *d_count, void *stream);
SIMDUTF_B200_API int b200_utf16_length_from_utf8(const char *d_in, size_t len, uint64_t *h_count, void *stream);
SIMDUTF_B200_API int b200_host_utf16_length_from_utf8(const char *h_in, size_t len, uint64_t *h_count);

/* ------------------------------------------------------------------------- */
/* UTF-8 -> UTF-16LE — implementation::convert_utf8_to_utf16le_with_errors     */
/* (:3743-3745), ::convert_utf8_to_utf16le (:3709, = count or 0 on error),     */
/* ::convert_valid_utf8_to_utf16le (:3815).                                    */
/* Semantics src/scalar/utf8_to_utf16/utf8_to_utf16.h:128-255.                 */
/* result = {SUCCESS, units written} or {error, input byte index}.  d_out must */
/* hold utf16_length_from_utf8(in) units (it is never overrun, even for        */
/* invalid input: reference tests/convert_utf8_to_utf16le_tests.cpp:23-52).    */
/* ------------------------------------------------------------------------- */
SIMDUTF_B200_API int b200_convert_utf8_to_utf16le_async(const char *d_in, size_t len, uint16_t *d_out, b200_result *d_res, void *stream);
SIMDUTF_B200_API int b200_convert_utf8_to_utf16le(const char *d_in, size_t len, uint16_t *d_out, b200_result *h_res, void *stream);
SIMDUTF_B200_API int b200_host_convert_utf8_to_utf16le(const char *h_in, size_t len, uint16_t *h_out, b200_result *h_res);

/* UTF-8 -> UTF-32 — implementation::convert_utf8_to_utf32_with_errors (:3799-3800),
 * ::convert_utf8_to_utf32 (:3781).  Semantics src/scalar/utf8_to_utf32/utf8_to_utf32.h:106-212.
 * d_out must hold count_utf8(in) words. */
SIMDUTF_B200_API int b200_convert_utf8_to_utf32_async(const char *d_in, size_t len, uint32_t *d_out, b200_result *d_res, void *stream);
SIMDUTF_B200_API int b200_convert_utf8_to_utf32(const char *d_in, size_t len, uint32_t *d_out, b200_result *h_res, void *stream);
SIMDUTF_B200_API int b200_host_convert_utf8_to_utf32(const char *h_in, size_t len, uint32_t *h_out, b200_result *h_res);

/* ------------------------------------------------------------------------- */
/* UTF-16LE counting — implementation::count_utf16le (:4767) ==                */
/* utf32_length_from_utf16le; ::utf8_length_from_utf16le (:4277-4278).         */
/* Semantics src/scalar/utf16.h:69-94.  `len` in 16-bit units.                 */
/* ------------------------------------------------------------------------- */
SIMDUTF_B200_API int b200_count_utf16le_async(const uint16_t *d_in, size_t len, uint64_t *d_count, void *stream);
SIMDUTF_B200_API int b200_count_utf16le(const uint16_t *d_in, size_t len, uint64_t *h_count, void *stream);
SIMDUTF_B200_API int b200_host_count_utf16le(const uint16_t *h_in, size_t len, uint64_t *h_count);
SIMDUTF_B200_API int b200_utf8_length_from_utf16le_async(const uint16_t *d_in, size_t len, uint64_t *d_count, void *stream);
SIMDUTF_B200_API int b200_utf8_length_from_utf16le(const uint16_t *d_in, size_t len, uint64_t *h_count, void *stream);
SIMDUTF_B200_API int b200_host_utf8_length_from_utf16le(const uint16_t *h_in, size_t len, uint64_t *h_count);

/* UTF-16LE validation — implementation::validate_utf16le_with_errors
 * (reference include/simdutf/implementation.h:3481-3483; src/scalar/utf16.h:39-67). */
SIMDUTF_B200_API int b200_validate_utf16le_with_errors_async(const uint16_t *d_in, size_t len, b200_result *d_res, void *stream);
SIMDUTF_B200_API int b200_validate_utf16le_with_errors(const uint16_t *d_in, size_t len, b200_result *h_res, void *stream);
SIMDUTF_B200_API int b200_host_validate_utf16le_with_errors(const uint16_t *h_in, size_t len, b200_result *h_res);

/* ------------------------------------------------------------------------- */
/* UTF-16LE -> UTF-8 — implementation::convert_utf16le_to_utf8_with_errors     */
/* (:4079-4080), ::convert_utf16le_to_utf8 (:4038).                            */
/* Semantics src/scalar/utf16_to_utf8/utf16_to_utf8.h:82-153.                  */
/* result = {SUCCESS, bytes written} or {SURROGATE, unit index}.  d_out must   */
/* hold utf8_length_from_utf16le(in) bytes.                                    */
/* ------------------------------------------------------------------------- */
SIMDUTF_B200_API int b200_convert_utf16le_to_utf8_async(const uint16_t *d_in, size_t len, char *d_out, b200_result *d_res, void *stream);
SIMDUTF_B200_API int b200_convert_utf16le_to_utf8(const uint16_t *d_in, size_t len, char *d_out, b200_result *h_res, void *stream);
SIMDUTF_B200_API int b200_host_convert_utf16le_to_utf8(const uint16_t *h_in, size_t len, char *h_out, b200_result *h_res);

/* ------------------------------------------------------------------------- */
/* UTF-16BE twins: the same kernels reading / writing   */
/* big-endian units.  implementation::convert_utf8_to_utf16be[_with_errors]    */
/* (reference include/simdutf/implementation.h:3727-3760), ::count_utf16be     */
/* (:4783) == utf32_length_from_utf16be, ::utf8_length_from_utf16be (:4299),   */
/* ::validate_utf16be_with_errors (:3499), ::convert_utf16be_to_utf8           */
/* [_with_errors] (:4058-4101), ::change_endianness_utf16 (:4567-4584; its     */
/* b200_result is {SUCCESS, len}).  Same result conventions as the LE entries. */
/* ------------------------------------------------------------------------- */
SIMDUTF_B200_API int b200_convert_utf8_to_utf16be_async(const char *d_in, size_t len, uint16_t *d_out, b200_result *d_res, void *stream);
SIMDUTF_B200_API int b200_convert_utf8_to_utf16be(const char *d_in, size_t len, uint16_t *d_out, b200_result *h_res, void *stream);
SIMDUTF_B200_API int b200_host_convert_utf8_to_utf16be(const char *h_in, size_t len, uint16_t *h_out, b200_result *h_res);
SIMDUTF_B200_API int b200_count_utf16be_async(const uint16_t *d_in, size_t len, uint64_t *d_count, void *stream);
SIMDUTF_B200_API int b200_count_utf16be(const uint16_t *d_in, size_t len, uint64_t *h_count, void *stream);
SIMDUTF_B200_API int b200_host_count_utf16be(const uint16_t *h_in, size_t len, uint64_t *h_count);
SIMDUTF_B200_API int b200_utf8_length_from_utf16be_async(const uint16_t *d_in, size_t len, uint64_t *d_count, void *stream);
SIMDUTF_B200_API int b200_utf8_length_from_utf16be(const uint16_t *d_in, size_t len, uint64_t *h_count, void *stream);
SIMDUTF_B200_API int b200_host_utf8_length_from_utf16be(const uint16_t *h_in, size_t len, uint64_t *h_count);
SIMDUTF_B200_API int b200_validate_utf16be_with_errors_async(const uint16_t *d_in, size_t len, b200_result *d_res, void *stream);
SIMDUTF_B200_API int b200_validate_utf16be_with_errors(const uint16_t *d_in, size_t len, b200_result *h_res, void *stream);
SIMDUTF_B200_API int b200_host_validate_utf16be_with_errors(const uint16_t *h_in, size_t len, b200_result *h_res);
SIMDUTF_B200_API int b200_convert_utf16be_to_utf8_async(const uint16_t *d_in, size_t len, char *d_out, b200_result *d_res, void *stream);
SIMDUTF_B200_API int b200_convert_utf16be_to_utf8(const uint16_t *d_in, size_t len, char *d_out, b200_result *h_res, void *stream);
SIMDUTF_B200_API int b200_host_convert_utf16be_to_utf8(const uint16_t *h_in, size_t len, char *h_out, b200_result *h_res);
SIMDUTF_B200_API int b200_change_endianness_utf16_async(const uint16_t *d_in, size_t len, uint16_t *d_out, b200_result *d_res, void *stream);
SIMDUTF_B200_API int b200_change_endianness_utf16(const uint16_t *d_in, size_t len, uint16_t *d_out, b200_result *h_res, void *stream);
SIMDUTF_B200_API int b200_host_change_endianness_utf16(const uint16_t *h_in, size_t len, uint16_t *h_out, b200_result *h_res);

/* ------------------------------------------------------------------------- */
/* WHATWG forgiving base64 decode — implementation::base64_to_binary_details   */
/* (:4902-4906) and ::base64_to_binary (:4866-4870; result derived from the    */
/* full_result as in include/simdutf/error.h:66-73).  Semantics                */
/* src/generic/base64.h:40-246 + src/scalar/base64.h:33-216.                   */
/* d_out must hold maximal_binary_length_from_base64(in) bytes.                */
/* ------------------------------------------------------------------------- */
SIMDUTF_B200_API int b200_base64_to_binary_async(const char *d_in, size_t len, char *d_out, uint64_t options, uint64_t last_chunk,
                                b200_full_result *d_res, void *stream);
SIMDUTF_B200_API int b200_base64_to_binary(const char *d_in, size_t len, char *d_out, uint64_t options, uint64_t last_chunk,
                          b200_full_result *h_res, void *stream);
SIMDUTF_B200_API int b200_host_base64_to_binary(const char *h_in, size_t len, char *h_out, uint64_t options, uint64_t last_chunk,
                               b200_full_result *h_res);
/* ------------------------------------------------------------------------- */
/* UTF-32 family (second part).  `len` in input elements. */
/* implementation::validate_utf32[_with_errors] (reference                      */
/* include/simdutf/implementation.h:3533-3569; src/scalar/utf32.h:10-38),       */
/* ::utf8_length_from_utf32 / ::utf16_length_from_utf32 (:4586-4620;            */
/* src/scalar/utf32.h:40-66), ::convert_utf32_to_utf8[_with_errors] (:4370-4410;*/
/* src/scalar/utf32_to_utf8/utf32_to_utf8.h:63-124),                            */
/* ::convert_utf32_to_utf16le/be[_with_errors] (:4440-4565;                     */
/* src/scalar/utf32_to_utf16/utf32_to_utf16.h:40-86),                           */
/* ::convert_utf16le/be_to_utf32[_with_errors] (:4103-4259;                     */
/* src/scalar/utf16_to_utf32/utf16_to_utf32.h:45-76).                           */
/* result = {SUCCESS, elements written} or {SURROGATE | TOO_LARGE, input index}.*/
/* ------------------------------------------------------------------------- */
SIMDUTF_B200_API int b200_validate_utf32_with_errors_async(const uint32_t *d_in, size_t len, b200_result *d_res, void *stream);
SIMDUTF_B200_API int b200_validate_utf32_with_errors(const uint32_t *d_in, size_t len, b200_result *h_res, void *stream);
SIMDUTF_B200_API int b200_host_validate_utf32_with_errors(const uint32_t *h_in, size_t len, b200_result *h_res);
SIMDUTF_B200_API int b200_utf8_length_from_utf32_async(const uint32_t *d_in, size_t len, uint64_t *d_res, void *stream);
SIMDUTF_B200_API int b200_utf8_length_from_utf32(const uint32_t *d_in, size_t len, uint64_t *h_res, void *stream);
SIMDUTF_B200_API int b200_host_utf8_length_from_utf32(const uint32_t *h_in, size_t len, uint64_t *h_res);
SIMDUTF_B200_API int b200_utf16_length_from_utf32_async(const uint32_t *d_in, size_t len, uint64_t *d_res, void *stream);
SIMDUTF_B200_API int b200_utf16_length_from_utf32(const uint32_t *d_in, size_t len, uint64_t *h_res, void *stream);
SIMDUTF_B200_API int b200_host_utf16_length_from_utf32(const uint32_t *h_in, size_t len, uint64_t *h_res);
SIMDUTF_B200_API int b200_convert_utf32_to_utf8_async(const uint32_t *d_in, size_t len, char *d_out, b200_result *d_res, void *stream);
SIMDUTF_B200_API int b200_convert_utf32_to_utf8(const uint32_t *d_in, size_t len, char *d_out, b200_result *h_res, void *stream);
SIMDUTF_B200_API int b200_host_convert_utf32_to_utf8(const uint32_t *h_in, size_t len, char *h_out, b200_result *h_res);
SIMDUTF_B200_API int b200_convert_utf32_to_utf16le_async(const uint32_t *d_in, size_t len, uint16_t *d_out, b200_result *d_res, void *stream);
SIMDUTF_B200_API int b200_convert_utf32_to_utf16le(const uint32_t *d_in, size_t len, uint16_t *d_out, b200_result *h_res, void *stream);
SIMDUTF_B200_API int b200_host_convert_utf32_to_utf16le(const uint32_t *h_in, size_t len, uint16_t *h_out, b200_result *h_res);
SIMDUTF_B200_API int b200_convert_utf32_to_utf16be_async(const uint32_t *d_in, size_t len, uint16_t *d_out, b200_result *d_res, void *stream);
SIMDUTF_B200_API int b200_convert_utf32_to_utf16be(const uint32_t *d_in, size_t len, uint16_t *d_out, b200_result *h_res, void *stream);
SIMDUTF_B200_API int b200_host_convert_utf32_to_utf16be(const uint32_t *h_in, size_t len, uint16_t *h_out, b200_result *h_res);
SIMDUTF_B200_API int b200_convert_utf16le_to_utf32_async(const uint16_t *d_in, size_t len, uint32_t *d_out, b200_result *d_res, void *stream);
SIMDUTF_B200_API int b200_convert_utf16le_to_utf32(const uint16_t *d_in, size_t len, uint32_t *d_out, b200_result *h_res, void *stream);
SIMDUTF_B200_API int b200_host_convert_utf16le_to_utf32(const uint16_t *h_in, size_t len, uint32_t *h_out, b200_result *h_res);
SIMDUTF_B200_API int b200_convert_utf16be_to_utf32_async(const uint16_t *d_in, size_t len, uint32_t *d_out, b200_result *d_res, void *stream);
SIMDUTF_B200_API int b200_convert_utf16be_to_utf32(const uint16_t *d_in, size_t len, uint32_t *d_out, b200_result *h_res, void *stream);
SIMDUTF_B200_API int b200_host_convert_utf16be_to_utf32(const uint16_t *h_in, size_t len, uint32_t *h_out, b200_result *h_res);
/* ------------------------------------------------------------------------- */
/* Latin-1 / ASCII family.  `len` in input elements.     */
/* implementation::validate_ascii[_with_errors] (reference                      */
/* include/simdutf/implementation.h:3409-3425; src/scalar/ascii.h:36-64),       */
/* ::utf8_length_from_latin1 (:4622-4634; src/scalar/latin1.h:9-19),            */
/* ::convert_latin1_to_utf8 / _utf16le / _utf16be / _utf32 (:3583-3692;         */
/* src/scalar/latin1_to_utf8/latin1_to_utf8.h:9-46),                            */
/* ::convert_utf8_to_latin1[_with_errors] (:3694-3760;                          */
/* src/scalar/utf8_to_latin1/utf8_to_latin1.h:83-149),                          */
/* ::convert_utf16le/be_to_latin1[_with_errors] (:3885-4019;                    */
/* src/scalar/utf16_to_latin1/utf16_to_latin1.h:38-92),                         */
/* ::convert_utf32_to_latin1[_with_errors] (:4299-4368;                         */
/* src/scalar/utf32_to_latin1/utf32_to_latin1.h:33-62).                         */
/* latin1_length_from_utf8 == count_utf8; the other latin1 length queries are    */
/* the identity and have no entry point.                                         */
/* result = {SUCCESS, elements written} or {error, input index}: TOO_LARGE for   */
/* anything above U+00FF; UTF-8 input also TOO_SHORT / TOO_LONG / OVERLONG /     */
/* HEADER_BITS exactly as the reference's scalar walk reports them.              */
/* ------------------------------------------------------------------------- */
SIMDUTF_B200_API int b200_validate_ascii_with_errors_async(const char *d_in, size_t len, b200_result *d_res, void *stream);
SIMDUTF_B200_API int b200_validate_ascii_with_errors(const char *d_in, size_t len, b200_result *h_res, void *stream);
SIMDUTF_B200_API int b200_host_validate_ascii_with_errors(const char *h_in, size_t len, b200_result *h_res);
SIMDUTF_B200_API int b200_utf8_length_from_latin1_async(const char *d_in, size_t len, uint64_t *d_res, void *stream);
SIMDUTF_B200_API int b200_utf8_length_from_latin1(const char *d_in, size_t len, uint64_t *h_res, void *stream);
SIMDUTF_B200_API int b200_host_utf8_length_from_latin1(const char *h_in, size_t len, uint64_t *h_res);
SIMDUTF_B200_API int b200_convert_latin1_to_utf8_async(const char *d_in, size_t len, char *d_out, b200_result *d_res, void *stream);
SIMDUTF_B200_API int b200_convert_latin1_to_utf8(const char *d_in, size_t len, char *d_out, b200_result *h_res, void *stream);
SIMDUTF_B200_API int b200_host_convert_latin1_to_utf8(const char *h_in, size_t len, char *h_out, b200_result *h_res);
SIMDUTF_B200_API int b200_convert_latin1_to_utf16le_async(const char *d_in, size_t len, uint16_t *d_out, b200_result *d_res, void *stream);
SIMDUTF_B200_API int b200_convert_latin1_to_utf16le(const char *d_in, size_t len, uint16_t *d_out, b200_result *h_res, void *stream);
SIMDUTF_B200_API int b200_host_convert_latin1_to_utf16le(const char *h_in, size_t len, uint16_t *h_out, b200_result *h_res);
SIMDUTF_B200_API int b200_convert_latin1_to_utf16be_async(const char *d_in, size_t len, uint16_t *d_out, b200_result *d_res, void *stream);
SIMDUTF_B200_API int b200_convert_latin1_to_utf16be(const char *d_in, size_t len, uint16_t *d_out, b200_result *h_res, void *stream);
SIMDUTF_B200_API int b200_host_convert_latin1_to_utf16be(const char *h_in, size_t len, uint16_t *h_out, b200_result *h_res);
SIMDUTF_B200_API int b200_convert_latin1_to_utf32_async(const char *d_in, size_t len, uint32_t *d_out, b200_result *d_res, void *stream);
SIMDUTF_B200_API int b200_convert_latin1_to_utf32(const char *d_in, size_t len, uint32_t *d_out, b200_result *h_res, void *stream);
SIMDUTF_B200_API int b200_host_convert_latin1_to_utf32(const char *h_in, size_t len, uint32_t *h_out, b200_result *h_res);
SIMDUTF_B200_API int b200_convert_utf8_to_latin1_async(const char *d_in, size_t len, char *d_out, b200_result *d_res, void *stream);
SIMDUTF_B200_API int b200_convert_utf8_to_latin1(const char *d_in, size_t len, char *d_out, b200_result *h_res, void *stream);
SIMDUTF_B200_API int b200_host_convert_utf8_to_latin1(const char *h_in, size_t len, char *h_out, b200_result *h_res);
SIMDUTF_B200_API int b200_convert_utf16le_to_latin1_async(const uint16_t *d_in, size_t len, char *d_out, b200_result *d_res, void *stream);
SIMDUTF_B200_API int b200_convert_utf16le_to_latin1(const uint16_t *d_in, size_t len, char *d_out, b200_result *h_res, void *stream);
SIMDUTF_B200_API int b200_host_convert_utf16le_to_latin1(const uint16_t *h_in, size_t len, char *h_out, b200_result *h_res);
SIMDUTF_B200_API int b200_convert_utf16be_to_latin1_async(const uint16_t *d_in, size_t len, char *d_out, b200_result *d_res, void *stream);
SIMDUTF_B200_API int b200_convert_utf16be_to_latin1(const uint16_t *d_in, size_t len, char *d_out, b200_result *h_res, void *stream);
SIMDUTF_B200_API int b200_host_convert_utf16be_to_latin1(const uint16_t *h_in, size_t len, char *h_out, b200_result *h_res);
SIMDUTF_B200_API int b200_convert_utf32_to_latin1_async(const uint32_t *d_in, size_t len, char *d_out, b200_result *d_res, void *stream);
SIMDUTF_B200_API int b200_convert_utf32_to_latin1(const uint32_t *d_in, size_t len, char *d_out, b200_result *h_res, void *stream);
SIMDUTF_B200_API int b200_host_convert_utf32_to_latin1(const uint32_t *h_in, size_t len, char *h_out, b200_result *h_res);
/* ------------------------------------------------------------------------- */
/* SURVEY.md §8f rank 4.                                                        */
/* implementation::to_well_formed_utf16le/be (reference                          */
/* include/simdutf/implementation.h:3498-3531; src/scalar/utf16.h:141-166):      */
/* out[i] = U+FFFD where in[i] is a lone surrogate, else in[i]; d_in == d_out is */
/* allowed; result = {SUCCESS, len}.                                             */
/* implementation::detect_encodings (:3344-3354;                                 */
/* src/fallback/implementation.cpp:8-32, src/encoding_types.cpp:32-49): the      */
/* encoding_type of a BOM if there is one, else the OR of UTF8 (1), UTF16_LE (2) */
/* and UTF32_LE (8) for every encoding the buffer validates as.  The device      */
/* flavours need a 4-byte aligned d_in.                                          */
/* ------------------------------------------------------------------------- */
SIMDUTF_B200_API int b200_to_well_formed_utf16le_async(const uint16_t *d_in, size_t len, uint16_t *d_out, b200_result *d_res, void *stream);
SIMDUTF_B200_API int b200_to_well_formed_utf16le(const uint16_t *d_in, size_t len, uint16_t *d_out, b200_result *h_res, void *stream);
SIMDUTF_B200_API int b200_host_to_well_formed_utf16le(const uint16_t *h_in, size_t len, uint16_t *h_out, b200_result *h_res);
SIMDUTF_B200_API int b200_to_well_formed_utf16be_async(const uint16_t *d_in, size_t len, uint16_t *d_out, b200_result *d_res, void *stream);
SIMDUTF_B200_API int b200_to_well_formed_utf16be(const uint16_t *d_in, size_t len, uint16_t *d_out, b200_result *h_res, void *stream);
SIMDUTF_B200_API int b200_host_to_well_formed_utf16be(const uint16_t *h_in, size_t len, uint16_t *h_out, b200_result *h_res);
SIMDUTF_B200_API int b200_detect_encodings_async(const char *d_in, size_t len, uint64_t *d_res, void *stream);
SIMDUTF_B200_API int b200_detect_encodings(const char *d_in, size_t len, uint64_t *h_res, void *stream);
SIMDUTF_B200_API int b200_host_detect_encodings(const char *h_in, size_t len, uint64_t *h_res);

/* base64 decode from char16_t input — implementation::base64_to_binary[_details](const char16_t*, ...)
 * (reference include/simdutf/implementation.h:4922-4939, 4976-5014): units above 0xFF are invalid characters
 * (src/scalar/base64.h:24-31, :125); everything else as for `char` input.  `len` in 16-bit units. */
SIMDUTF_B200_API int b200_base64_to_binary_utf16_async(const uint16_t *d_in, size_t len, char *d_out, uint64_t options, uint64_t last_chunk,
                                      b200_full_result *d_res, void *stream);
SIMDUTF_B200_API int b200_base64_to_binary_utf16(const uint16_t *d_in, size_t len, char *d_out, uint64_t options, uint64_t last_chunk,
                                b200_full_result *h_res, void *stream);
SIMDUTF_B200_API int b200_host_base64_to_binary_utf16(const uint16_t *h_in, size_t len, char *h_out, uint64_t options, uint64_t last_chunk,
                                     b200_full_result *h_res);
/* binary_to_base64 — implementation::binary_to_base64 (reference
 * include/simdutf/implementation.h:4941-4960; semantics src/scalar/base64.h:435-491).  options: base64_default (0),
 * base64_url (1), and the reverse-padding variants (2, 3).  d_out must hold base64_length_from_binary(len, options)
 * characters (src/scalar/base64.h:515-533); result = {SUCCESS, characters written}. */
SIMDUTF_B200_API int b200_binary_to_base64_async(const char *d_in, size_t len, char *d_out, uint64_t options, b200_result *d_res, void *stream);
SIMDUTF_B200_API int b200_binary_to_base64(const char *d_in, size_t len, char *d_out, uint64_t options, b200_result *h_res, void *stream);
SIMDUTF_B200_API int b200_host_binary_to_base64(const char *h_in, size_t len, char *h_out, uint64_t options, b200_result *h_res);
SIMDUTF_B200_API size_t b200_base64_length_from_binary(size_t len, uint64_t options);
/* implementation::maximal_binary_length_from_base64 (reference src/implementation.cpp:87-90 ->
 * src/scalar/base64.h:493-513): O(1), looks at the last two characters only; host pointer. */
SIMDUTF_B200_API size_t b200_host_maximal_binary_length_from_base64(const char *h_in, size_t len);

/* ------------------------------------------------------------------------- */
/* Shard-cut helpers (multi-GPU / chunked streaming): the largest prefix that   */
/* does not end inside a character — simdutf::trim_partial_utf8 / _utf16le      */
/* (reference src/implementation.cpp:2502-2525, src/scalar/utf8.h:257-288,      */
/*  src/scalar/utf16.h:114-124; split idiom benchmarks/threaded.cpp:69-74).     */
/* Host pointers; they read at most the last 3 bytes / 1 unit.                  */
/* ------------------------------------------------------------------------- */
SIMDUTF_B200_API size_t b200_host_trim_partial_utf8(const char *h_in, size_t len);
SIMDUTF_B200_API size_t b200_host_trim_partial_utf16le(const uint16_t *h_in, size_t len);
/* Same, on device data (synchronous, copies <= 4 bytes back). */
SIMDUTF_B200_API int b200_trim_partial_utf8(const char *d_in, size_t len, size_t *h_trimmed, void *stream);

/* ------------------------------------------------------------------------- */
/* Sharded (multi-GPU) path, SURVEY.md §8e: a buffer cut with the helpers above */
/* is processed shard by shard by the single-GPU entry points; every shard      */
/* leaves a triplet {input length, b200_result} (3 x uint64, the result written */
/* in place by the *_async call).  After the triplets of all shards have been   */
/* gathered in shard order (one NCCL all_gather across processes),              */
/* b200_sharded_combine_async turns them into the global result — the first     */
/* error in buffer order, i.e. the minimum of (position << 8 | code), the value */
/* an NCCL min-allreduce of that key yields, or {SUCCESS, total} — and this     */
/* shard's global input / output offsets, on the device, in one launch.         */
/* count_is_length != 0: the operation's success count is the validated input   */
/* length (validate_*) rather than an output size.                              */
/* ------------------------------------------------------------------------- */
typedef struct b200_sharded_result {
  int32_t error;        /* global simdutf::error_code */
  uint32_t reserved_;
  uint64_t count;       /* total output elements on success, global input position on error */
  uint64_t in_offset;   /* this shard's first input element in the whole buffer */
  uint64_t out_offset;  /* this shard's first output element in the whole output */
} b200_sharded_result;
SIMDUTF_B200_API int b200_sharded_combine_async(const uint64_t *d_gathered, int world, int rank, int count_is_length,
                                                b200_sharded_result *d_out, void *stream);

/* One process, several devices (SURVEY.md §8b "_mgpu variants taking per-device shard descriptors", §7 step 5).  */
/* Shard i is device-resident on shards[i].device: d_in / len (code units) / d_out as for the single-GPU entry     */
/* point of the same name (d_out unused by validate / length).  Every shard runs on its own device and stream; the */
/* triplets are exchanged with one ncclAllGather over NVLink (communicators created once per device list with      */
/* ncclCommInitAll; NCCL is resolved with dlopen("libnccl.so.2"), and 24-byte peer copies carry the triplets when  */
/* it is absent or when two shards share a device); the combine kernel runs on every device.  h_results[i] is the  */
/* global result as device i sees it (all equal) plus shard i's offsets.  The call returns when all devices are    */
/* done.  For b200_mgpu_utf16_length_from_utf8 the global count is the sum of the shard lengths in UTF-16 units.   */
/* b200_mgpu_last_gather(): 1 if the last call exchanged over NCCL, 2 if it used peer copies.                      */
typedef struct b200_shard {
  int32_t device;
  uint32_t reserved_;
  const void *d_in;
  uint64_t len;
  void *d_out;
} b200_shard;
SIMDUTF_B200_API int b200_mgpu_validate_utf8_with_errors(const b200_shard *shards, int n, b200_sharded_result *h_results);
SIMDUTF_B200_API int b200_mgpu_utf16_length_from_utf8(const b200_shard *shards, int n, b200_sharded_result *h_results);
SIMDUTF_B200_API int b200_mgpu_convert_utf8_to_utf16le(const b200_shard *shards, int n, b200_sharded_result *h_results);
SIMDUTF_B200_API int b200_mgpu_convert_utf8_to_utf32(const b200_shard *shards, int n, b200_sharded_result *h_results);
SIMDUTF_B200_API int b200_mgpu_convert_utf16le_to_utf8(const b200_shard *shards, int n, b200_sharded_result *h_results);
SIMDUTF_B200_API int b200_mgpu_last_gather(void);

/* Host-pointer calls over several devices: after b200_host_set_devices(n) the b200_host_* calls of the calling      */
/* thread — hence the C++ virtuals of simdutf::b200::implementation — cut large buffers into segments and deal them  */
/* round-robin to n devices (b200_get_device(), +1, ...), each with its own staging ring, streams and PCIe link; the */
/* results are folded exactly like shards.  Default 1.  Small inputs always use one device.                          */
SIMDUTF_B200_API int b200_host_set_devices(int n);
SIMDUTF_B200_API int b200_host_get_devices(void);
/* ------------------------------------------------------------------------- */
/* Many small strings per launch ("callers at the right   */
/* granularity"; the reference's callers of this shape: tools/sutf.cpp:131-336,  */
/* the per-string loops of tests/validate_utf8_with_errors_tests.cpp:54-69).     */
/* A single-string call costs a launch and a synchronisation however short the  */
/* string; a batch pays them once.  String i is d_data[d_offsets[i] ..           */
/* d_offsets[i + 1]) — one packed buffer and n + 1 offsets, the layout of an      */
/* Arrow string column — and gets the result the single-string entry point of    */
/* the same name gives (implementation::validate_utf8_with_errors :3396,         */
/* count_utf8 :4802, utf16_length_from_utf8 :3863, convert_utf8_to_utf16le/be    */
/* _with_errors :3727-3760).  convert: string i's units start at                 */
/* d_out[d_out_offsets[i]], or, with d_out_offsets == NULL, at d_out[d_offsets[i]]*/
/* (a string never yields more units than it has bytes, so a d_out of            */
/* d_offsets[n] units needs no length pass); d_results[i].count is the number    */
/* of units.  The host flavour packs the strings into a pinned window itself.    */
/* ------------------------------------------------------------------------- */
SIMDUTF_B200_API int b200_validate_utf8_batch_async(const char *d_data, const uint64_t *d_offsets, size_t n,
                                                    b200_result *d_results, void *stream);
SIMDUTF_B200_API int b200_count_utf8_batch_async(const char *d_data, const uint64_t *d_offsets, size_t n,
                                                 uint64_t *d_counts, void *stream);
SIMDUTF_B200_API int b200_utf16_length_from_utf8_batch_async(const char *d_data, const uint64_t *d_offsets, size_t n,
                                                             uint64_t *d_counts, void *stream);
SIMDUTF_B200_API int b200_convert_utf8_to_utf16le_batch_async(const char *d_data, const uint64_t *d_offsets, size_t n,
                                                              uint16_t *d_out, const uint64_t *d_out_offsets,
                                                              b200_result *d_results, void *stream);
SIMDUTF_B200_API int b200_convert_utf8_to_utf16be_batch_async(const char *d_data, const uint64_t *d_offsets, size_t n,
                                                              uint16_t *d_out, const uint64_t *d_out_offsets,
                                                              b200_result *d_results, void *stream);
SIMDUTF_B200_API int b200_host_validate_utf8_batch(const char *const *h_strings, const size_t *h_lens, size_t n,
                                                   b200_result *h_results);
SIMDUTF_B200_API int b200_host_count_utf8_batch(const char *const *h_strings, const size_t *h_lens, size_t n,
                                                uint64_t *h_counts);
SIMDUTF_B200_API int b200_host_utf16_length_from_utf8_batch(const char *const *h_strings, const size_t *h_lens, size_t n,
                                                            uint64_t *h_counts);
/* h_outs[i] must hold utf16_length_from_utf8(string i) units; it is written only when string i converts without error */
SIMDUTF_B200_API int b200_host_convert_utf8_to_utf16le_batch(const char *const *h_strings, const size_t *h_lens, size_t n,
                                                             uint16_t *const *h_outs, b200_result *h_results);

/* Experiment knobs for tools/ and profiles/ ("segment_mb", "no_nccl", "conv_variant", "dbg_lo", "dbg_hi"; 0 = default). */
/* The library never reads the environment.                                                                          */
SIMDUTF_B200_API int b200_set_tuning(const char *name, int value);

#ifdef __cplusplus
}
#endif
#endif /* SIMDUTF_B200_H */
