timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "utf16be or utf16_random or config3 or utf8_small" 2>&1 | tail -15
for t in convert_utf8_to_utf16be_tests convert_utf8_to_utf16be_with_errors_tests convert_valid_utf8_to_utf16be_tests convert_utf16be_to_utf8_tests convert_utf16be_to_utf8_with_errors_tests convert_valid_utf16be_to_utf8_tests count_utf16be validate_utf16be_basic_tests validate_utf16be_with_errors_tests utf8_length_from_utf16_tests; do
  timeout 300 simdutf_b200/build/with_b200/$t -a b200 > gpurun_out/ref_$t.log 2>&1; echo "$t rc=$? $(grep -c OK gpurun_out/ref_$t.log) OK"
done
