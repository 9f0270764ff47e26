python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
for t in convert_utf8_to_utf16le_tests validate_utf8_with_errors_tests convert_utf32_to_utf8_tests convert_latin1_to_utf8_tests bele_tests select_implementation; do timeout 200 simdutf_b200/build/with_b200/$t -a b200 > gpurun_out/ref_$t.log 2>&1; echo "$t rc=$? OK=$(grep -c ' OK' gpurun_out/ref_$t.log)"; done
