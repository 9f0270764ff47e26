"""Generates, at BUILD time and into a git-ignored directory, the glue that plugs `simdutf::b200::implementation`
into the UNMODIFIED reference tree.  Nothing generated here is committed: the reference's
sources are read where they lie (/root/reference) and only derived files are written.

    python tools/gen_b200_cxx.py <reference_root> <out_dir>

Writes into <out_dir>:
  b200_decls.inc       `override` declarations for EVERY pure virtual of class simdutf::implementation
                       (reference include/simdutf/implementation.h:3302-5066), parsed from the header itself.
  b200_stubs.inc       definitions for the pure virtuals OUTSIDE the hot path: they return the same
                       "unsupported" values as the reference's unsupported_implementation
                       (src/implementation.cpp:792-1253): 0 / false / result(OTHER, 0).  (No CPU fallback is
                       allowed inside the b200 implementation, so forwarding to another kernel is not an option.)
  implementation_b200.cpp   the reference's src/implementation.cpp with four insertions: the b200 header, a
                       get_b200_singleton(), one list entry AFTER the fallback singleton (so automatic
                       detection never picks it; it is selected by name, by assigning
                       get_active_implementation(), or with SIMDUTF_FORCE_IMPLEMENTATION=b200), and the
                       SIMDUTF_IMPLEMENTATION_B200 term in the SIMDUTF_SINGLE_IMPLEMENTATION sum.
  simdutf_b200_unity.cpp    the reference's unity TU src/simdutf.cpp with its `#include "implementation.cpp"`
                       redirected to the file above.
"""
import os
import re
import sys

ref, out = sys.argv[1], sys.argv[2]
os.makedirs(out, exist_ok=True)

# (name, first parameter type) of the virtuals written by hand in simdutf_b200/csrc/b200_implementation.cpp
HOT = {
    ("validate_utf8", "const char *"), ("validate_utf8_with_errors", "const char *"),
    ("count_utf8", "const char *"), ("utf16_length_from_utf8", "const char *"), ("utf32_length_from_utf8", "const char *"),
    ("convert_utf8_to_utf16le", "const char *"), ("convert_utf8_to_utf16le_with_errors", "const char *"),
    ("convert_valid_utf8_to_utf16le", "const char *"),
    ("convert_utf8_to_utf32", "const char *"), ("convert_utf8_to_utf32_with_errors", "const char *"),
    ("convert_valid_utf8_to_utf32", "const char *"),
    ("count_utf16le", "const char16_t *"), ("utf8_length_from_utf16le", "const char16_t *"),
    ("utf32_length_from_utf16le", "const char16_t *"),
    ("validate_utf16le", "const char16_t *"), ("validate_utf16le_with_errors", "const char16_t *"),
    ("convert_utf16le_to_utf8", "const char16_t *"), ("convert_utf16le_to_utf8_with_errors", "const char16_t *"),
    ("convert_valid_utf16le_to_utf8", "const char16_t *"),
    ("base64_to_binary", "const char *"), ("base64_to_binary_details", "const char *"),
    # UTF-16BE twins + change_endianness_utf16
    ("count_utf16be", "const char16_t *"), ("utf32_length_from_utf16be", "const char16_t *"),
    ("utf8_length_from_utf16be", "const char16_t *"),
    ("validate_utf16be", "const char16_t *"), ("validate_utf16be_with_errors", "const char16_t *"),
    ("convert_utf8_to_utf16be", "const char *"), ("convert_utf8_to_utf16be_with_errors", "const char *"),
    ("convert_valid_utf8_to_utf16be", "const char *"),
    ("convert_utf16be_to_utf8", "const char16_t *"), ("convert_utf16be_to_utf8_with_errors", "const char16_t *"),
    ("convert_valid_utf16be_to_utf8", "const char16_t *"),
    ("change_endianness_utf16", "const char16_t *"),
    ("binary_to_base64", "const char *"),
    ("base64_to_binary", "const char16_t *"), ("base64_to_binary_details", "const char16_t *"),
    # UTF-32 family (second part)
    ("validate_utf32", "const char32_t *"), ("validate_utf32_with_errors", "const char32_t *"),
    ("utf8_length_from_utf32", "const char32_t *"), ("utf16_length_from_utf32", "const char32_t *"),
    ("convert_utf32_to_utf8", "const char32_t *"), ("convert_utf32_to_utf8_with_errors", "const char32_t *"),
    ("convert_valid_utf32_to_utf8", "const char32_t *"),
    ("convert_utf32_to_utf16le", "const char32_t *"), ("convert_utf32_to_utf16le_with_errors", "const char32_t *"),
    ("convert_valid_utf32_to_utf16le", "const char32_t *"),
    ("convert_utf32_to_utf16be", "const char32_t *"), ("convert_utf32_to_utf16be_with_errors", "const char32_t *"),
    ("convert_valid_utf32_to_utf16be", "const char32_t *"),
    ("convert_utf16le_to_utf32", "const char16_t *"), ("convert_utf16le_to_utf32_with_errors", "const char16_t *"),
    ("convert_valid_utf16le_to_utf32", "const char16_t *"),
    ("convert_utf16be_to_utf32", "const char16_t *"), ("convert_utf16be_to_utf32_with_errors", "const char16_t *"),
    ("convert_valid_utf16be_to_utf32", "const char16_t *"),
    # Latin-1 / ASCII family
    ("validate_ascii", "const char *"), ("validate_ascii_with_errors", "const char *"),
    ("utf8_length_from_latin1", "const char *"), ("latin1_length_from_utf8", "const char *"),
    ("convert_latin1_to_utf8", "const char *"), ("convert_latin1_to_utf16le", "const char *"),
    ("convert_latin1_to_utf16be", "const char *"), ("convert_latin1_to_utf32", "const char *"),
    ("convert_utf8_to_latin1", "const char *"), ("convert_utf8_to_latin1_with_errors", "const char *"),
    ("convert_valid_utf8_to_latin1", "const char *"),
    ("convert_utf16le_to_latin1", "const char16_t *"), ("convert_utf16le_to_latin1_with_errors", "const char16_t *"),
    ("convert_valid_utf16le_to_latin1", "const char16_t *"),
    ("convert_utf16be_to_latin1", "const char16_t *"), ("convert_utf16be_to_latin1_with_errors", "const char16_t *"),
    ("convert_valid_utf16be_to_latin1", "const char16_t *"),
    ("convert_utf32_to_latin1", "const char32_t *"), ("convert_utf32_to_latin1_with_errors", "const char32_t *"),
    ("convert_valid_utf32_to_latin1", "const char32_t *"),
    # SURVEY.md §8f rank 4
    ("to_well_formed_utf16le", "const char16_t *"), ("to_well_formed_utf16be", "const char16_t *"),
    ("detect_encodings", "const char *"),
}

hdr = open(os.path.join(ref, "include/simdutf/implementation.h")).read()
body = hdr[hdr.index("class implementation {"):]
pure = re.findall(r"virtual\s+([\w:<>\s\*&]+?)\s*\b(\w+)\s*\(([^)]*)\)\s*const\s+noexcept\s*=\s*0\s*;", body, flags=re.S)


def norm(s):
    return " ".join(s.split())


def first_type(args):
    a = norm(args).split(",")[0]
    m = re.match(r"(.*?[\*&]?)\s*\w+$", a)
    return norm(m.group(1)) if m else a


def strip_names(args):
    """parameter list without parameter names (silences unused-parameter warnings in the stubs)"""
    outp = []
    for a in norm(args).split(","):
        a = a.split("=")[0].strip()  # drop default arguments
        if not a:
            continue
        m = re.match(r"(.*?[\*&\s])\s*(\w+)$", a)
        outp.append(norm(m.group(1)) if m else a)
    return ", ".join(outp)


RET = {"size_t": "return 0;", "bool": "return false;", "void": "", "int": "return 0;",
       "result": "return result(error_code::OTHER, 0);", "full_result": "return full_result(error_code::OTHER, 0, 0);"}
decls, stubs, seen_hot = [], [], set()
for ret, name, args in pure:
    ret = norm(ret)
    decls.append(f"  {ret} {name}({norm(args)}) const noexcept override;")
    key = (name, first_type(args))
    if key in HOT:
        seen_hot.add(key)
        continue
    stubs.append(f"{ret} implementation::{name}({strip_names(args)}) const noexcept {{ {RET[ret]} }}")
missing = HOT - seen_hot
assert not missing, f"hot-path virtuals not found in the reference header: {missing}"
open(os.path.join(out, "b200_decls.inc"), "w").write("\n".join(decls) + "\n")
open(os.path.join(out, "b200_stubs.inc"), "w").write("\n".join(stubs) + "\n")

# --- registry patch (three insertions, checked) -----------------------------------------------------------------
impl = open(os.path.join(ref, "src/implementation.cpp")).read()
anchor1 = "#if SIMDUTF_IMPLEMENTATION_FALLBACK\nstatic const fallback::implementation *get_fallback_singleton() {"
anchor2 = "#if SIMDUTF_IMPLEMENTATION_FALLBACK\n          get_fallback_singleton(),\n#endif\n"
assert impl.count(anchor1) == 1 and impl.count(anchor2) == 1, "reference registry layout changed"
impl = impl.replace(anchor1, '''static const b200::implementation *get_b200_singleton() {
  static const b200::implementation b200_singleton{};
  return &b200_singleton;
}
''' + anchor1)
impl = impl.replace(anchor2, anchor2 + "          get_b200_singleton(),\n")
# SIMDUTF_SINGLE_IMPLEMENTATION (reference src/implementation.cpp:107-112, SURVEY gotcha G3): with exactly one CPU kernel
# compiled in, the free functions bypass the active-implementation pointer and b200 could never be selected; count it.
anchor3 = "SIMDUTF_IMPLEMENTATION_LASX + SIMDUTF_IMPLEMENTATION_FALLBACK =="
assert impl.count(anchor3) == 1, "reference SIMDUTF_SINGLE_IMPLEMENTATION layout changed"
impl = impl.replace(anchor3, "SIMDUTF_IMPLEMENTATION_LASX + SIMDUTF_IMPLEMENTATION_FALLBACK + SIMDUTF_IMPLEMENTATION_B200 ==")
impl = "#ifndef SIMDUTF_IMPLEMENTATION_B200\n#define SIMDUTF_IMPLEMENTATION_B200 1\n#endif\n" + impl
impl = '#include "b200_implementation.h"\n' + impl
open(os.path.join(out, "implementation_b200.cpp"), "w").write(impl)

unity = open(os.path.join(ref, "src/simdutf.cpp")).read()
assert unity.count('#include "implementation.cpp"') == 1
unity = unity.replace('#include "implementation.cpp"', '#include "implementation_b200.cpp"')
open(os.path.join(out, "simdutf_b200_unity.cpp"), "w").write(unity)
print(f"{len(pure)} pure virtuals: {len(seen_hot)} hot-path (hand-written), {len(stubs)} stubs -> {out}")
