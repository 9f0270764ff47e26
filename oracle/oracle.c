/* oracle/oracle.c — TEST INFRASTRUCTURE ONLY (see oracle.h for the rules).
 *
 * CPU restatement of the reference's scalar semantics for the hot path
 *  and the families around it (§8f ranks 1-4: UTF-16BE, UTF-32,
 * Latin-1 / ASCII, base64 encode, to_well_formed_utf16, detect_encodings).  It
 * is written from the behaviour described in SURVEY.md Appendix A and the cited
 * reference lines, as one decode-one-character routine per encoding rather
 * than a transcription of the reference's loops.
 *
 * PARITY PINNED: every function here is checked against the reference's golden
 * vectors / recorded outputs (tests/golden/golden.json, golden_next.json) and,
 * where oracle/_ref exists, differentially against the unmodified reference
 * library's icelake / haswell / fallback kernels (tests/test_oracle.py).
 */
#include "oracle.h"
#include <string.h>

/* ------------------------------------------------------------------------- */
/* UTF-8: decode exactly one character at in[pos].                           */
/* Follows reference src/scalar/utf8.h:102-200 (validate_with_errors) and    */
/* src/scalar/utf8_to_utf16/utf8_to_utf16.h:128-255 (same error rules, same  */
/* order: length/continuation checks -> OVERLONG -> SURROGATE / TOO_LARGE).  */
/* Returns the error code; on success *cp and *adv are set.                  */
/* ------------------------------------------------------------------------- */
static int is_cont(uint8_t b) { return (b & 0xC0) == 0x80; }

static int utf8_decode_one(const uint8_t *in, size_t len, size_t pos, uint32_t *cp, unsigned *adv) {
  uint8_t b0 = in[pos];
  if (b0 < 0x80) { *cp = b0; *adv = 1; return ORACLE_SUCCESS; }
  if ((b0 & 0xE0) == 0xC0) {
    if (pos + 2 > len || !is_cont(in[pos + 1])) return ORACLE_TOO_SHORT;
    uint32_t c = ((uint32_t)(b0 & 0x1F) << 6) | (in[pos + 1] & 0x3F);
    if (c < 0x80) return ORACLE_OVERLONG;
    *cp = c; *adv = 2; return ORACLE_SUCCESS;
  }
  if ((b0 & 0xF0) == 0xE0) {
    if (pos + 3 > len || !is_cont(in[pos + 1]) || !is_cont(in[pos + 2])) return ORACLE_TOO_SHORT;
    uint32_t c = ((uint32_t)(b0 & 0x0F) << 12) | ((uint32_t)(in[pos + 1] & 0x3F) << 6) | (in[pos + 2] & 0x3F);
    if (c < 0x800) return ORACLE_OVERLONG;
    if (c >= 0xD800 && c <= 0xDFFF) return ORACLE_SURROGATE;
    *cp = c; *adv = 3; return ORACLE_SUCCESS;
  }
  if ((b0 & 0xF8) == 0xF0) {
    if (pos + 4 > len || !is_cont(in[pos + 1]) || !is_cont(in[pos + 2]) || !is_cont(in[pos + 3]))
      return ORACLE_TOO_SHORT;
    uint32_t c = ((uint32_t)(b0 & 0x07) << 18) | ((uint32_t)(in[pos + 1] & 0x3F) << 12) |
                 ((uint32_t)(in[pos + 2] & 0x3F) << 6) | (in[pos + 3] & 0x3F);
    if (c <= 0xFFFF) return ORACLE_OVERLONG;
    if (c > 0x10FFFF) return ORACLE_TOO_LARGE;
    *cp = c; *adv = 4; return ORACLE_SUCCESS;
  }
  /* stray continuation byte, or a header with 5+ leading ones (0xF8..0xFF) */
  return is_cont(b0) ? ORACLE_TOO_LONG : ORACLE_HEADER_BITS;
}

/* reference src/scalar/utf8.h:102-200; empty input -> {SUCCESS,0}
 * (src/icelake/implementation.cpp:204-206). */
oracle_result oracle_validate_utf8_with_errors(const uint8_t *in, size_t len) {
  oracle_result r;
  size_t pos = 0;
  while (pos < len) {
    uint32_t cp; unsigned adv;
    int e = utf8_decode_one(in, len, pos, &cp, &adv);
    if (e) { r.error = e; r.count = pos; return r; }
    pos += adv;
  }
  r.error = ORACLE_SUCCESS; r.count = len;
  return r;
}

/* reference include/simdutf/implementation.h:3378-3379 (validate_utf8) */
int oracle_validate_utf8(const uint8_t *in, size_t len) {
  return oracle_validate_utf8_with_errors(in, len).error == ORACLE_SUCCESS;
}

/* reference src/scalar/utf8.h:230-241 — never validates */
uint64_t oracle_count_utf8(const uint8_t *in, size_t len) {
  uint64_t n = 0;
  for (size_t i = 0; i < len; i++) n += !is_cont(in[i]);
  return n;
}

/* reference src/scalar/utf8.h:243-255 — never validates */
uint64_t oracle_utf16_length_from_utf8(const uint8_t *in, size_t len) {
  uint64_t n = 0;
  for (size_t i = 0; i < len; i++) n += (uint64_t)!is_cont(in[i]) + (in[i] >= 0xF0);
  return n;
}

/* reference src/icelake/implementation.cpp:1596-1599 (== count_utf8) */
uint64_t oracle_utf32_length_from_utf8(const uint8_t *in, size_t len) { return oracle_count_utf8(in, len); }

/* reference src/scalar/utf8_to_utf16/utf8_to_utf16.h:128-255 (LITTLE endian; host is LE) */
oracle_result oracle_convert_utf8_to_utf16le_with_errors(const uint8_t *in, size_t len, uint16_t *out) {
  oracle_result r;
  size_t pos = 0; uint64_t w = 0;
  while (pos < len) {
    uint32_t cp; unsigned adv;
    int e = utf8_decode_one(in, len, pos, &cp, &adv);
    if (e) { r.error = e; r.count = pos; return r; }
    if (cp < 0x10000) {
      out[w++] = (uint16_t)cp;
    } else {
      cp -= 0x10000;
      out[w++] = (uint16_t)(0xD800 + (cp >> 10));
      out[w++] = (uint16_t)(0xDC00 + (cp & 0x3FF));
    }
    pos += adv;
  }
  r.error = ORACLE_SUCCESS; r.count = w;
  return r;
}

/* reference src/scalar/utf8_to_utf16/utf8_to_utf16.h:9-126: units written, 0 on any error */
uint64_t oracle_convert_utf8_to_utf16le(const uint8_t *in, size_t len, uint16_t *out) {
  oracle_result r = oracle_convert_utf8_to_utf16le_with_errors(in, len, out);
  return r.error ? 0 : r.count;
}

/* reference src/scalar/utf8_to_utf32/utf8_to_utf32.h:106-212 */
oracle_result oracle_convert_utf8_to_utf32_with_errors(const uint8_t *in, size_t len, uint32_t *out) {
  oracle_result r;
  size_t pos = 0; uint64_t w = 0;
  while (pos < len) {
    uint32_t cp; unsigned adv;
    int e = utf8_decode_one(in, len, pos, &cp, &adv);
    if (e) { r.error = e; r.count = pos; return r; }
    out[w++] = cp;
    pos += adv;
  }
  r.error = ORACLE_SUCCESS; r.count = w;
  return r;
}

uint64_t oracle_convert_utf8_to_utf32(const uint8_t *in, size_t len, uint32_t *out) {
  oracle_result r = oracle_convert_utf8_to_utf32_with_errors(in, len, out);
  return r.error ? 0 : r.count;
}

/* ------------------------------------------------------------------------- */
/* UTF-16 (LE and BE)                                                         */
/* The reference's scalar routines are templates on the endianness that read  */
/* every unit through `!match_system(big_endian) ? u16_swap_bytes(w) : w`     */
/* (src/scalar/utf16.h:44, :73, :85; src/scalar/utf16_to_utf8/utf16_to_utf8.h */
/* :93; src/scalar/utf8_to_utf16/utf8_to_utf16.h:143-146 for the stores); the */
/* host is little-endian, so `be` means "swap".                               */
/* ------------------------------------------------------------------------- */
static int is_high(uint16_t w) { return (w & 0xFC00) == 0xD800; }
static int is_low(uint16_t w) { return (w & 0xFC00) == 0xDC00; }
static uint16_t swap16(uint16_t w) { return (uint16_t)((w >> 8) | (w << 8)); }
static uint16_t ld16(const uint16_t *in, size_t i, int be) { return be ? swap16(in[i]) : in[i]; }

/* reference src/scalar/utf16.h:39-67 */
static oracle_result validate_utf16_impl(const uint16_t *in, size_t len, int be) {
  oracle_result r;
  size_t pos = 0;
  while (pos < len) {
    uint16_t w = ld16(in, pos, be);
    if ((w & 0xF800) == 0xD800) {
      if (!is_high(w) || pos + 1 >= len || !is_low(ld16(in, pos + 1, be))) { r.error = ORACLE_SURROGATE; r.count = pos; return r; }
      pos += 2;
    } else {
      pos += 1;
    }
  }
  r.error = ORACLE_SUCCESS; r.count = len;
  return r;
}
oracle_result oracle_validate_utf16le_with_errors(const uint16_t *in, size_t len) { return validate_utf16_impl(in, len, 0); }
oracle_result oracle_validate_utf16be_with_errors(const uint16_t *in, size_t len) { return validate_utf16_impl(in, len, 1); }

/* reference src/scalar/utf16.h:69-78 */
static uint64_t count_utf16_impl(const uint16_t *in, size_t len, int be) {
  uint64_t n = 0;
  for (size_t i = 0; i < len; i++) n += !is_low(ld16(in, i, be));
  return n;
}
uint64_t oracle_count_utf16le(const uint16_t *in, size_t len) { return count_utf16_impl(in, len, 0); }
uint64_t oracle_count_utf16be(const uint16_t *in, size_t len) { return count_utf16_impl(in, len, 1); }

/* reference src/scalar/utf16.h:80-94 — every surrogate unit counts 2 */
static uint64_t utf8_length_from_utf16_impl(const uint16_t *in, size_t len, int be) {
  uint64_t n = 0;
  for (size_t i = 0; i < len; i++) {
    uint16_t w = ld16(in, i, be);
    n += 1 + (w > 0x7F) + ((w > 0x7FF && w <= 0xD7FF) || w >= 0xE000);
  }
  return n;
}
uint64_t oracle_utf8_length_from_utf16le(const uint16_t *in, size_t len) { return utf8_length_from_utf16_impl(in, len, 0); }
uint64_t oracle_utf8_length_from_utf16be(const uint16_t *in, size_t len) { return utf8_length_from_utf16_impl(in, len, 1); }

/* reference src/scalar/utf16.h:96-105 */
uint64_t oracle_utf32_length_from_utf16le(const uint16_t *in, size_t len) { return oracle_count_utf16le(in, len); }
uint64_t oracle_utf32_length_from_utf16be(const uint16_t *in, size_t len) { return oracle_count_utf16be(in, len); }

/* reference src/scalar/utf16_to_utf8/utf16_to_utf8.h:82-153 */
static oracle_result convert_utf16_to_utf8_impl(const uint16_t *in, size_t len, uint8_t *out, int be) {
  oracle_result r;
  size_t pos = 0; uint64_t w = 0;
  while (pos < len) {
    uint32_t u = ld16(in, pos, be);
    if (u < 0x80) {
      out[w++] = (uint8_t)u; pos++;
    } else if (u < 0x800) {
      out[w++] = (uint8_t)(0xC0 | (u >> 6));
      out[w++] = (uint8_t)(0x80 | (u & 0x3F)); pos++;
    } else if ((u & 0xF800) != 0xD800) {
      out[w++] = (uint8_t)(0xE0 | (u >> 12));
      out[w++] = (uint8_t)(0x80 | ((u >> 6) & 0x3F));
      out[w++] = (uint8_t)(0x80 | (u & 0x3F)); pos++;
    } else {
      if (pos + 1 >= len || !is_high((uint16_t)u) || !is_low(ld16(in, pos + 1, be))) {
        r.error = ORACLE_SURROGATE; r.count = pos; return r;
      }
      uint32_t cp = 0x10000 + ((u - 0xD800) << 10) + (ld16(in, pos + 1, be) - 0xDC00u);
      out[w++] = (uint8_t)(0xF0 | (cp >> 18));
      out[w++] = (uint8_t)(0x80 | ((cp >> 12) & 0x3F));
      out[w++] = (uint8_t)(0x80 | ((cp >> 6) & 0x3F));
      out[w++] = (uint8_t)(0x80 | (cp & 0x3F)); pos += 2;
    }
  }
  r.error = ORACLE_SUCCESS; r.count = w;
  return r;
}
oracle_result oracle_convert_utf16le_to_utf8_with_errors(const uint16_t *in, size_t len, uint8_t *out) {
  return convert_utf16_to_utf8_impl(in, len, out, 0);
}
oracle_result oracle_convert_utf16be_to_utf8_with_errors(const uint16_t *in, size_t len, uint8_t *out) {
  return convert_utf16_to_utf8_impl(in, len, out, 1);
}

/* reference src/scalar/utf16_to_utf8/utf16_to_utf8.h:9-80: bytes written, 0 on error */
uint64_t oracle_convert_utf16le_to_utf8(const uint16_t *in, size_t len, uint8_t *out) {
  oracle_result r = oracle_convert_utf16le_to_utf8_with_errors(in, len, out);
  return r.error ? 0 : r.count;
}

/* reference src/scalar/utf8_to_utf16/utf8_to_utf16.h:128-255 with big_endian stores (:143-146, :186-189, :228-235):
 * the same units, byte-swapped.  On error the units written so far are swapped too. */
oracle_result oracle_convert_utf8_to_utf16be_with_errors(const uint8_t *in, size_t len, uint16_t *out) {
  /* units written before the error position = utf16 length of the valid prefix */
  oracle_result r = oracle_convert_utf8_to_utf16le_with_errors(in, len, out);
  uint64_t n = r.error ? oracle_utf16_length_from_utf8(in, r.count) : r.count;
  for (uint64_t i = 0; i < n; i++) out[i] = swap16(out[i]);
  return r;
}

/* reference src/scalar/utf16.h:107-112 (change_endianness_utf16) */
void oracle_change_endianness_utf16(const uint16_t *in, size_t len, uint16_t *out) {
  for (size_t i = 0; i < len; i++) out[i] = swap16(in[i]);
}

/* ------------------------------------------------------------------------- */
/* UTF-32 (second part)                                 */
/* ------------------------------------------------------------------------- */
/* reference src/scalar/utf32.h:24-38 */
oracle_result oracle_validate_utf32_with_errors(const uint32_t *in, size_t len) {
  oracle_result r;
  for (size_t pos = 0; pos < len; pos++) {
    uint32_t w = in[pos];
    if (w > 0x10FFFF) { r.error = ORACLE_TOO_LARGE; r.count = pos; return r; }
    if (w >= 0xD800 && w <= 0xDFFF) { r.error = ORACLE_SURROGATE; r.count = pos; return r; }
  }
  r.error = ORACLE_SUCCESS; r.count = len;
  return r;
}
/* reference src/scalar/utf32.h:40-53 */
uint64_t oracle_utf8_length_from_utf32(const uint32_t *in, size_t len) {
  uint64_t n = 0;
  for (size_t i = 0; i < len; i++) n += 1 + (in[i] > 0x7F) + (in[i] > 0x7FF) + (in[i] > 0xFFFF);
  return n;
}
/* reference src/scalar/utf32.h:55-66 */
uint64_t oracle_utf16_length_from_utf32(const uint32_t *in, size_t len) {
  uint64_t n = 0;
  for (size_t i = 0; i < len; i++) n += 1 + (in[i] > 0xFFFF);
  return n;
}
/* reference src/scalar/utf32_to_utf8/utf32_to_utf8.h:63-124 */
oracle_result oracle_convert_utf32_to_utf8_with_errors(const uint32_t *in, size_t len, uint8_t *out) {
  oracle_result r;
  uint64_t w = 0;
  for (size_t pos = 0; pos < len; pos++) {
    uint32_t c = in[pos];
    if ((c & 0xFFFFFF80) == 0) {
      out[w++] = (uint8_t)c;
    } else if ((c & 0xFFFFF800) == 0) {
      out[w++] = (uint8_t)((c >> 6) | 0xC0);
      out[w++] = (uint8_t)((c & 0x3F) | 0x80);
    } else if ((c & 0xFFFF0000) == 0) {
      if (c >= 0xD800 && c <= 0xDFFF) { r.error = ORACLE_SURROGATE; r.count = pos; return r; }
      out[w++] = (uint8_t)((c >> 12) | 0xE0);
      out[w++] = (uint8_t)(((c >> 6) & 0x3F) | 0x80);
      out[w++] = (uint8_t)((c & 0x3F) | 0x80);
    } else {
      if (c > 0x10FFFF) { r.error = ORACLE_TOO_LARGE; r.count = pos; return r; }
      out[w++] = (uint8_t)((c >> 18) | 0xF0);
      out[w++] = (uint8_t)(((c >> 12) & 0x3F) | 0x80);
      out[w++] = (uint8_t)(((c >> 6) & 0x3F) | 0x80);
      out[w++] = (uint8_t)((c & 0x3F) | 0x80);
    }
  }
  r.error = ORACLE_SUCCESS; r.count = w;
  return r;
}
/* reference src/scalar/utf32_to_utf16/utf32_to_utf16.h:40-86 */
static oracle_result convert_utf32_to_utf16_impl(const uint32_t *in, size_t len, uint16_t *out, int be) {
  oracle_result r;
  uint64_t w = 0;
  for (size_t pos = 0; pos < len; pos++) {
    uint32_t c = in[pos];
    if ((c & 0xFFFF0000) == 0) {
      if (c >= 0xD800 && c <= 0xDFFF) { r.error = ORACLE_SURROGATE; r.count = pos; return r; }
      out[w++] = be ? swap16((uint16_t)c) : (uint16_t)c;
    } else {
      if (c > 0x10FFFF) { r.error = ORACLE_TOO_LARGE; r.count = pos; return r; }
      c -= 0x10000;
      uint16_t hi = (uint16_t)(0xD800 + (c >> 10)), lo = (uint16_t)(0xDC00 + (c & 0x3FF));
      out[w++] = be ? swap16(hi) : hi;
      out[w++] = be ? swap16(lo) : lo;
    }
  }
  r.error = ORACLE_SUCCESS; r.count = w;
  return r;
}
oracle_result oracle_convert_utf32_to_utf16le_with_errors(const uint32_t *in, size_t len, uint16_t *out) {
  return convert_utf32_to_utf16_impl(in, len, out, 0);
}
oracle_result oracle_convert_utf32_to_utf16be_with_errors(const uint32_t *in, size_t len, uint16_t *out) {
  return convert_utf32_to_utf16_impl(in, len, out, 1);
}
/* reference src/scalar/utf16_to_utf32/utf16_to_utf32.h:45-76 */
static oracle_result convert_utf16_to_utf32_impl(const uint16_t *in, size_t len, uint32_t *out, int be) {
  oracle_result r;
  size_t pos = 0; uint64_t w = 0;
  while (pos < len) {
    uint16_t u = ld16(in, pos, be);
    if ((u & 0xF800) != 0xD800) {
      out[w++] = u; pos++;
    } else {
      uint16_t diff = (uint16_t)(u - 0xD800);
      if (diff > 0x3FF || pos + 1 >= len) { r.error = ORACLE_SURROGATE; r.count = pos; return r; }
      uint16_t diff2 = (uint16_t)(ld16(in, pos + 1, be) - 0xDC00);
      if (diff2 > 0x3FF) { r.error = ORACLE_SURROGATE; r.count = pos; return r; }
      out[w++] = ((uint32_t)diff << 10) + diff2 + 0x10000; pos += 2;
    }
  }
  r.error = ORACLE_SUCCESS; r.count = w;
  return r;
}
oracle_result oracle_convert_utf16le_to_utf32_with_errors(const uint16_t *in, size_t len, uint32_t *out) {
  return convert_utf16_to_utf32_impl(in, len, out, 0);
}
oracle_result oracle_convert_utf16be_to_utf32_with_errors(const uint16_t *in, size_t len, uint32_t *out) {
  return convert_utf16_to_utf32_impl(in, len, out, 1);
}

/* ------------------------------------------------------------------------- */
/* Latin-1 / ASCII                                     */
/* ------------------------------------------------------------------------- */
/* reference src/scalar/ascii.h:36-64 (the 16-byte skip does not change the answer) */
oracle_result oracle_validate_ascii_with_errors(const uint8_t *in, size_t len) {
  oracle_result r;
  for (size_t pos = 0; pos < len; pos++)
    if (in[pos] >= 0x80) { r.error = ORACLE_TOO_LARGE; r.count = pos; return r; }
  r.error = ORACLE_SUCCESS; r.count = len;
  return r;
}
/* reference src/scalar/latin1.h:9-19 */
uint64_t oracle_utf8_length_from_latin1(const uint8_t *in, size_t len) {
  uint64_t n = len;
  for (size_t i = 0; i < len; i++) n += in[i] >> 7;
  return n;
}
/* reference src/scalar/latin1_to_utf8/latin1_to_utf8.h:9-46 */
uint64_t oracle_convert_latin1_to_utf8(const uint8_t *in, size_t len, uint8_t *out) {
  uint64_t w = 0;
  for (size_t pos = 0; pos < len; pos++) {
    uint8_t b = in[pos];
    if ((b & 0x80) == 0) out[w++] = b;
    else { out[w++] = (uint8_t)((b >> 6) | 0xC0); out[w++] = (uint8_t)((b & 0x3F) | 0x80); }
  }
  return w;
}
/* reference src/scalar/latin1_to_utf16/latin1_to_utf16.h:10-24 */
uint64_t oracle_convert_latin1_to_utf16(const uint8_t *in, size_t len, uint16_t *out, int be) {
  for (size_t i = 0; i < len; i++) out[i] = be ? swap16((uint16_t)in[i]) : (uint16_t)in[i];
  return len;
}
/* reference src/scalar/latin1_to_utf32/latin1_to_utf32.h:10-18 */
uint64_t oracle_convert_latin1_to_utf32(const uint8_t *in, size_t len, uint32_t *out) {
  for (size_t i = 0; i < len; i++) out[i] = in[i];
  return len;
}
/* reference src/scalar/utf8_to_latin1/utf8_to_latin1.h:83-149 (the 16-byte ASCII skip does not change the answer) */
oracle_result oracle_convert_utf8_to_latin1_with_errors(const uint8_t *in, size_t len, uint8_t *out) {
  oracle_result r;
  size_t pos = 0; uint64_t w = 0;
  while (pos < len) {
    uint8_t lead = in[pos];
    if (lead < 0x80) { out[w++] = lead; pos++; }
    else if ((lead & 0xE0) == 0xC0) {
      if (pos + 1 >= len || (in[pos + 1] & 0xC0) != 0x80) { r.error = ORACLE_TOO_SHORT; r.count = pos; return r; }
      uint32_t cp = (uint32_t)(lead & 0x1F) << 6 | (in[pos + 1] & 0x3F);
      if (cp < 0x80) { r.error = ORACLE_OVERLONG; r.count = pos; return r; }
      if (cp > 0xFF) { r.error = ORACLE_TOO_LARGE; r.count = pos; return r; }
      out[w++] = (uint8_t)cp; pos += 2;
    } else if ((lead & 0xF0) == 0xE0 || (lead & 0xF8) == 0xF0) {
      r.error = ORACLE_TOO_LARGE; r.count = pos; return r;
    } else {
      r.error = (lead & 0xC0) == 0x80 ? ORACLE_TOO_LONG : ORACLE_HEADER_BITS; r.count = pos; return r;
    }
  }
  r.error = ORACLE_SUCCESS; r.count = w;
  return r;
}
/* reference src/scalar/utf16_to_latin1/utf16_to_latin1.h:38-92 */
oracle_result oracle_convert_utf16_to_latin1_with_errors(const uint16_t *in, size_t len, uint8_t *out, int be) {
  oracle_result r;
  for (size_t pos = 0; pos < len; pos++) {
    uint16_t u = ld16(in, pos, be);
    if (u & 0xFF00) { r.error = ORACLE_TOO_LARGE; r.count = pos; return r; }
    out[pos] = (uint8_t)u;
  }
  r.error = ORACLE_SUCCESS; r.count = len;
  return r;
}
/* reference src/scalar/utf32_to_latin1/utf32_to_latin1.h:33-62 */
oracle_result oracle_convert_utf32_to_latin1_with_errors(const uint32_t *in, size_t len, uint8_t *out) {
  oracle_result r;
  for (size_t pos = 0; pos < len; pos++) {
    if (in[pos] & 0xFFFFFF00u) { r.error = ORACLE_TOO_LARGE; r.count = pos; return r; }
    out[pos] = (uint8_t)in[pos];
  }
  r.error = ORACLE_SUCCESS; r.count = len;
  return r;
}

/* ------------------------------------------------------------------------- */
/* to_well_formed_utf16 and detect_encodings           */
/* ------------------------------------------------------------------------- */
/* reference src/scalar/utf16.h:141-166 */
void oracle_to_well_formed_utf16(const uint16_t *in, size_t len, uint16_t *out, int be) {
  const uint16_t repl = be ? 0xFDFF : 0xFFFD;
  int high_prev = 0;
  size_t i = 0;
  for (; i < len; i++) {
    uint16_t u = ld16(in, i, be);
    int high = (u & 0xFC00) == 0xD800, low = (u & 0xFC00) == 0xDC00;
    if (high_prev && !low) out[i - 1] = repl;
    out[i] = (!high_prev && low) ? repl : in[i];
    high_prev = high;
  }
  if (high_prev) out[i - 1] = repl;
}
/* reference src/encoding_types.cpp:32-49 (BOM::check_bom) and src/fallback/implementation.cpp:8-32 */
int oracle_detect_encodings(const uint8_t *in, size_t len) {
  if (len >= 2 && in[0] == 0xFF && in[1] == 0xFE) return (len >= 4 && in[2] == 0 && in[3] == 0) ? 8 : 2;
  if (len >= 2 && in[0] == 0xFE && in[1] == 0xFF) return 4;
  if (len >= 4 && in[0] == 0 && in[1] == 0 && in[2] == 0xFE && in[3] == 0xFF) return 16;
  if (len >= 4 && in[0] == 0xEF && in[1] == 0xBB && in[2] == 0xBF) return 1;
  int out = 0;
  if (oracle_validate_utf8_with_errors(in, len).error == ORACLE_SUCCESS) out |= 1;
  if (len % 2 == 0 && oracle_validate_utf16le_with_errors((const uint16_t *)in, len / 2).error == ORACLE_SUCCESS) out |= 2;
  if (len % 4 == 0 && oracle_validate_utf32_with_errors((const uint32_t *)in, len / 4).error == ORACLE_SUCCESS) out |= 8;
  return out;
}

/* ------------------------------------------------------------------------- */
/* Base64 (WHATWG forgiving decode)                                          */
/* ------------------------------------------------------------------------- */
/* Character class: 0..63 sextet, 64 = ASCII whitespace (' ' \t \n \r \f),   */
/* 255 = anything else.  Mirrors the three lookup tables at reference        */
/* src/tables/base64_tables.h:791-849 (default / url / default_or_url).     */
static uint8_t b64_class(uint8_t c, uint64_t options) {
  int url = (options & ORACLE_B64_URL) != 0;
  int both = (options & ORACLE_B64_DEFAULT_OR_URL) != 0;
  if (c >= 'A' && c <= 'Z') return (uint8_t)(c - 'A');
  if (c >= 'a' && c <= 'z') return (uint8_t)(c - 'a' + 26);
  if (c >= '0' && c <= '9') return (uint8_t)(c - '0' + 52);
  if (c == ' ' || c == '\t' || c == '\n' || c == '\r' || c == '\f') return 64;
  if (c == '+' && (both || !url)) return 62;
  if (c == '/' && (both || !url)) return 63;
  if (c == '-' && (both || url)) return 62;
  if (c == '_' && (both || url)) return 63;
  return 255;
}

static int b64_ignore_garbage(uint64_t options) {
  return options == ORACLE_B64_URL_ACCEPT_GARBAGE || options == ORACLE_B64_DEFAULT_ACCEPT_GARBAGE ||
         options == ORACLE_B64_DEFAULT_OR_URL_ACCEPT_GARBAGE;
}

/* reference src/scalar/base64.h:493-513 */
uint64_t oracle_maximal_binary_length_from_base64(const uint8_t *in, size_t len) {
  size_t padding = 0;
  if (len > 0 && in[len - 1] == '=') {
    padding++;
    if (len > 1 && in[len - 2] == '=') padding++;
  }
  size_t actual = len - padding;
  if (actual % 4 <= 1) return actual / 4 * 3;
  return actual / 4 * 3 + (actual % 4) - 1;
}

/* The decode proper: reference src/fallback/implementation.cpp:570-621
 * (trailing whitespace / '=' stripping, empty-input rules, padding
 * consistency check) around reference src/scalar/base64.h:33-216
 * (base64_tail_decode: sextet gathering, last-chunk rules).  Only
 * (error, input_count) are pinned on INVALID_BASE64_CHARACTER — output_count
 * differs between the reference's own kernels there; this
 * oracle reports the scalar/fallback value.
 */
oracle_full_result oracle_base64_to_binary_details(const uint8_t *in, size_t len, uint8_t *out,
                                                   uint64_t options, uint64_t last_chunk) {
  oracle_full_result r;
  const int garbage = b64_ignore_garbage(options);
  /* Strip trailing whitespace and up to two '=' (whitespace allowed between).
   * The SIMD kernels (generic/base64.h:50-73, icelake_base64.inl.cpp) skip this
   * step entirely in accept_garbage mode, whereas the fallback kernel strips
   * regardless (fallback/implementation.cpp:577-596); the two differ only in
   * full_result.input_count on success.  North star = icelake/haswell, so the
   * SIMD behaviour is the one restated here. */
  while (!garbage && len > 0 && b64_class(in[len - 1], options) == 64) len--;
  size_t equallocation = len;
  size_t equalsigns = 0;
  if (!garbage && len > 0 && in[len - 1] == '=') {
    equallocation = len - 1; len--; equalsigns = 1;
    while (len > 0 && b64_class(in[len - 1], options) == 64) len--;
    if (len > 0 && in[len - 1] == '=') { equallocation = len - 1; len--; equalsigns = 2; }
  }
  if (len == 0) { /* fallback/implementation.cpp:597-608 */
    r.input_count = 0; r.output_count = 0; r.error = ORACLE_SUCCESS;
    if (!garbage && equalsigns > 0) {
      if (last_chunk == ORACLE_STRICT) { r.error = ORACLE_BASE64_INPUT_REMAINDER; }
      else if (last_chunk == ORACLE_STOP_BEFORE_PARTIAL) { r.error = ORACLE_SUCCESS; }
      else { r.error = ORACLE_INVALID_BASE64_CHARACTER; r.input_count = equallocation; }
    }
    return r;
  }
  /* scalar/base64.h:76-216: gather sextets four at a time */
  size_t src = 0; uint64_t dst = 0;
  for (;;) {
    uint8_t q[4]; unsigned idx = 0;
    size_t chunk_start = src;
    while (idx < 4 && src < len) {
      uint8_t v = b64_class(in[src], options);
      if (v <= 63) { q[idx++] = v; }
      else if (!garbage && v > 64) {
        r.error = ORACLE_INVALID_BASE64_CHARACTER; r.input_count = src; r.output_count = dst; return r;
      }
      src++;
    }
    if (idx == 4) {
      uint32_t t = ((uint32_t)q[0] << 18) | ((uint32_t)q[1] << 12) | ((uint32_t)q[2] << 6) | q[3];
      out[dst++] = (uint8_t)(t >> 16); out[dst++] = (uint8_t)(t >> 8); out[dst++] = (uint8_t)t;
      continue;
    }
    /* partial (or empty) final chunk: scalar/base64.h:138-200 */
    if (!garbage && last_chunk == ORACLE_STRICT && idx != 1 && ((idx + equalsigns) & 3) != 0) {
      r.error = ORACLE_BASE64_INPUT_REMAINDER; r.input_count = src; r.output_count = dst; return r;
    }
    if (!garbage && last_chunk == ORACLE_STOP_BEFORE_PARTIAL && ((idx + equalsigns) & 3) != 0) {
      src = chunk_start;
      while (src < len && b64_class(in[src], options) > 63) src++;
      r.error = ORACLE_SUCCESS; r.input_count = src; r.output_count = dst; return r;
    }
    if (idx == 2) {
      uint32_t t = ((uint32_t)q[0] << 18) | ((uint32_t)q[1] << 12);
      if (!garbage && last_chunk == ORACLE_STRICT && (t & 0xFFFF)) {
        r.error = ORACLE_BASE64_EXTRA_BITS; r.input_count = src; r.output_count = dst; return r;
      }
      out[dst++] = (uint8_t)(t >> 16);
    } else if (idx == 3) {
      uint32_t t = ((uint32_t)q[0] << 18) | ((uint32_t)q[1] << 12) | ((uint32_t)q[2] << 6);
      if (!garbage && last_chunk == ORACLE_STRICT && (t & 0xFF)) {
        r.error = ORACLE_BASE64_EXTRA_BITS; r.input_count = src; r.output_count = dst; return r;
      }
      out[dst++] = (uint8_t)(t >> 16); out[dst++] = (uint8_t)(t >> 8);
    } else if (!garbage && idx == 1 && last_chunk != ORACLE_STOP_BEFORE_PARTIAL) {
      r.error = ORACLE_BASE64_INPUT_REMAINDER; r.input_count = src; r.output_count = dst; return r;
    }
    r.error = ORACLE_SUCCESS; r.input_count = src; r.output_count = dst;
    break;
  }
  /* fallback/implementation.cpp:611-619: padding must complete the last quantum */
  if (last_chunk != ORACLE_STOP_BEFORE_PARTIAL && r.error == ORACLE_SUCCESS && equalsigns > 0 && !garbage) {
    if ((r.output_count % 3 == 0) || ((r.output_count % 3) + 1 + equalsigns != 4)) {
      r.error = ORACLE_INVALID_BASE64_CHARACTER; r.input_count = equallocation;
    }
  }
  return r;
}

/* full_result -> result: reference include/simdutf/error.h:66-73 */
oracle_result oracle_base64_to_binary(const uint8_t *in, size_t len, uint8_t *out, uint64_t options, uint64_t last_chunk) {
  oracle_full_result f = oracle_base64_to_binary_details(in, len, out, options, last_chunk);
  oracle_result r;
  r.error = f.error;
  r.count = (f.error == ORACLE_SUCCESS || f.error == ORACLE_BASE64_INPUT_REMAINDER) ? f.output_count : f.input_count;
  return r;
}

static int b64_use_padding(uint64_t options) {
  return ((options & ORACLE_B64_URL) == 0) ^ ((options & ORACLE_B64_REVERSE_PADDING) == ORACLE_B64_REVERSE_PADDING);
}

/* reference src/scalar/base64.h:515-533 */
uint64_t oracle_base64_length_from_binary(size_t len, uint64_t options) {
  if (!b64_use_padding(options)) return len / 3 * 4 + ((len % 3) ? (len % 3) + 1 : 0);
  return (len + 2) / 3 * 4;
}

/* reference src/scalar/base64.h:435-491 (encoder; used by tests to build inputs) */
uint64_t oracle_binary_to_base64(const uint8_t *in, size_t len, uint8_t *out, uint64_t options) {
  static const char std_abc[] = "ABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz0123456789+/";
  static const char url_abc[] = "ABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz0123456789-_";
  const char *abc = (options & ORACLE_B64_URL) ? url_abc : std_abc;
  const int pad = b64_use_padding(options);
  uint64_t w = 0; size_t i = 0;
  for (; i + 3 <= len; i += 3) {
    uint32_t t = ((uint32_t)in[i] << 16) | ((uint32_t)in[i + 1] << 8) | in[i + 2];
    out[w++] = (uint8_t)abc[t >> 18]; out[w++] = (uint8_t)abc[(t >> 12) & 63];
    out[w++] = (uint8_t)abc[(t >> 6) & 63]; out[w++] = (uint8_t)abc[t & 63];
  }
  if (len - i == 1) {
    uint32_t t = (uint32_t)in[i] << 16;
    out[w++] = (uint8_t)abc[t >> 18]; out[w++] = (uint8_t)abc[(t >> 12) & 63];
    if (pad) { out[w++] = '='; out[w++] = '='; }
  } else if (len - i == 2) {
    uint32_t t = ((uint32_t)in[i] << 16) | ((uint32_t)in[i + 1] << 8);
    out[w++] = (uint8_t)abc[t >> 18]; out[w++] = (uint8_t)abc[(t >> 12) & 63]; out[w++] = (uint8_t)abc[(t >> 6) & 63];
    if (pad) { out[w++] = '='; }
  }
  return w;
}

/* ------------------------------------------------------------------------- */
/* Shard-cut helpers                                                          */
/* ------------------------------------------------------------------------- */
/* reference src/scalar/utf8.h:257-288 (trim_partial_utf8) */
uint64_t oracle_trim_partial_utf8(const uint8_t *in, size_t len) {
  if (len >= 1 && in[len - 1] >= 0xC0) return len - 1;
  if (len >= 2 && in[len - 2] >= 0xE0) return len - 2;
  if (len >= 3 && in[len - 3] >= 0xF0) return len - 3;
  return len;
}

/* reference src/scalar/utf16.h:114-124 (trim_partial_utf16<LITTLE>) */
uint64_t oracle_trim_partial_utf16le(const uint16_t *in, size_t len) {
  if (len <= 1) return len;
  return len - (is_high(in[len - 1]) ? 1 : 0);
}
