timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "utf32_family" 2>&1 | tail -25
for op in utf32to8 utf32to16 utf32to16be utf16to32 validate32 len8from32 b64encode; do timeout 300 python tools/prof_one.py $op $((1<<30)) 5 2>&1 | tail -2; done
timeout 1500 python -m pytest tests/test_reference_suite.py -m gpu -q -k "utf32" 2>&1 | tail -15
for t in convert_utf16be_to_utf8_with_errors_tests convert_utf32_to_utf8_with_errors_tests convert_utf16le_to_utf32_with_errors_tests; do
  timeout 900 simdutf_b200/build/with_b200/$t -a b200 > gpurun_out/ref_$t.log 2>&1; echo "$t rc=$? OK=$(grep -c ' OK' gpurun_out/ref_$t.log)"; grep -v " OK" gpurun_out/ref_$t.log | head -6
done
