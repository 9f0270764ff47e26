"""GPU test (-m gpu): the reference's OWN test binaries, linked against the unmodified reference tree plus
simdutf::b200::implementation (simdutf_b200/build/with_b200, built by build.build_reference_integration() in the
build container — they travel to the GPU box as prebuilt files), run with `-a b200`: every TEST in them then
exercises the CUDA kernels through the C++ virtuals (reference tests/helpers/test.cpp:143-207).

Every pure virtual of simdutf::implementation is served by a kernel (SURVEY.md §8a + §8f ranks 1-4), so every
binary listed here must pass; SLOW holds the ones that are green but consist of 10^6..10^8 tiny calls."""
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
D = os.path.join(ROOT, "simdutf_b200", "build", "with_b200")

HOT = [
    "validate_utf8_basic_tests", "validate_utf8_puzzler_tests", "validate_utf8_brute_force_tests",
    "validate_utf8_with_errors_tests", "convert_utf8_to_utf16le_tests", "convert_utf8_to_utf16le_with_errors_tests",
    "convert_valid_utf8_to_utf16le_tests", "convert_utf8_to_utf32_tests", "convert_utf8_to_utf32_with_errors_tests",
    "convert_valid_utf8_to_utf32_tests", "convert_utf16le_to_utf8_tests", "convert_utf16le_to_utf8_with_errors_tests",
    "convert_valid_utf16le_to_utf8_tests", "count_utf8", "count_utf16le", "validate_utf16le_basic_tests",
    "validate_utf16le_with_errors_tests", "select_implementation",
    # UTF-16BE twins.  Not asserted here: convert_utf8_to_utf16be_with_errors_tests (green, but
    # its 100 trials per case take > 10 minutes of ~60 us host-path calls).
    "convert_utf8_to_utf16be_tests", "convert_valid_utf8_to_utf16be_tests",
    "convert_utf16be_to_utf8_tests", "convert_utf16be_to_utf8_with_errors_tests", "convert_valid_utf16be_to_utf8_tests",
    "count_utf16be", "validate_utf16be_basic_tests", "validate_utf16be_with_errors_tests", "utf8_length_from_utf16_tests",
    # base64 both ways, char and char16_t input, every option: the whole reference binary
    "base64_tests",
    # UTF-32 family (second part): every reference test binary of the family
    "validate_utf32_with_errors_tests",
    "convert_utf32_to_utf8_tests", "convert_utf32_to_utf8_with_errors_tests", "convert_valid_utf32_to_utf8_tests",
    "convert_utf32_to_utf16le_tests", "convert_utf32_to_utf16le_with_errors_tests", "convert_valid_utf32_to_utf16le_tests",
    "convert_utf32_to_utf16be_tests", "convert_utf32_to_utf16be_with_errors_tests", "convert_valid_utf32_to_utf16be_tests",
    "convert_utf16le_to_utf32_tests", "convert_utf16le_to_utf32_with_errors_tests", "convert_valid_utf16le_to_utf32_tests",
    "convert_utf16be_to_utf32_tests", "convert_utf16be_to_utf32_with_errors_tests", "convert_valid_utf16be_to_utf32_tests",
    # Latin-1 / ASCII family; the slow binaries of the family are in SLOW below
    "validate_ascii_basic_tests", "validate_ascii_with_errors_tests", "convert_latin1_to_utf8_tests",
    "convert_latin1_to_utf16le_tests", "convert_latin1_to_utf16be_tests", "convert_latin1_to_utf32_tests",
    "convert_utf8_to_latin1_tests", "convert_valid_utf8_to_latin1_tests",
    "convert_utf16le_to_latin1_tests", "convert_utf16le_to_latin1_tests_with_errors", "convert_valid_utf16le_to_latin1_tests",
    "convert_utf16be_to_latin1_tests", "convert_utf16be_to_latin1_tests_with_errors", "convert_valid_utf16be_to_latin1_tests",
    "convert_valid_utf32_to_latin1_tests",
    "bele_tests",
    # SURVEY.md §8f rank 4
    "to_well_formed_utf16_tests", "detect_encodings_tests",
    # whole-API binaries: null / empty arguments, the differential fuzzers (random_fuzzer walks EVERY supported
    # implementation, b200 included, and compares them), round trips over every family, the README's examples, the
    # std::span and atomic-ref front ends (free functions: routed to b200 with SIMDUTF_FORCE_IMPLEMENTATION, reference
    # src/implementation.cpp:1294-1305), and the internal_tests() hook (simdutf::b200::implementation provides two)
    "null_safety_tests", "random_fuzzer", "basic_fuzzer", "special_tests", "readme_tests", "span_tests",
    "atomic_base64_tests", "internal_tests",
]
# what a passing run must print besides exiting 0 (default: the harness's "OK")
MARKER = {"select_implementation": None, "random_fuzzer": "testing: b200", "internal_tests": "b200 host path over all devices"}

# Green as well, but made of 10^6..10^8 calls on <= 256-byte inputs (a CPU does those in nanoseconds, a host-path call
# costs ~30-60 us): minutes to half an hour each on the GPU box.  Run with B200_SLOW_TESTS=1.  Measured on a B200:
#   validate_utf32_basic_tests 112 s, convert_utf8_to_latin1_with_errors_tests 343 s, convert_utf32_to_latin1_tests
#   ~900 s (64 million calls; all passed); convert_utf8_to_utf16be_with_errors_tests > 600 s and
#   convert_utf32_to_latin1_with_errors_tests (64 million calls) were cut off with every finished case OK.
SLOW = [
    "validate_utf32_basic_tests", "convert_utf8_to_latin1_with_errors_tests", "convert_utf32_to_latin1_tests",
    "convert_utf8_to_utf16be_with_errors_tests", "convert_utf32_to_latin1_with_errors_tests",
]
_slow = pytest.mark.skipif(os.environ.get("B200_SLOW_TESTS") != "1", reason="minutes of tiny host calls; set B200_SLOW_TESTS=1")


@pytest.mark.parametrize("name", HOT + [pytest.param(n, marks=_slow) for n in SLOW])
def test_reference_binary_with_b200(name):
    exe = os.path.join(D, name)
    if not os.path.exists(exe):
        pytest.skip("reference test binaries were not built (no /root/reference at build time)")
    env = dict(os.environ, SIMDUTF_FORCE_IMPLEMENTATION="b200")  # free functions (simdutf::validate_utf8, ...) use b200 too
    p = subprocess.run([exe, "-a", "b200"], capture_output=True, text=True, timeout=3000 if name in SLOW else 600, env=env)
    out = p.stdout + p.stderr
    assert "unsupported by the current processor" not in out, "b200 reported itself unsupported on a GPU box"
    assert p.returncode == 0, out[-3000:]
    marker = MARKER.get(name, "OK")
    if marker is not None:
        assert marker in out, out[-3000:]
