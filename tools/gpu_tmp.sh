timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "well_formed" 2>&1 | tail -25
for t in convert_valid_utf32_to_latin1_tests bele_tests to_well_formed_utf16_tests detect_encodings_tests; do
  s=$(date +%s); timeout 400 simdutf_b200/build/with_b200/$t -a b200 > gpurun_out/ref_$t.log 2>&1; echo "$t rc=$? OK=$(grep -c ' OK' gpurun_out/ref_$t.log) secs=$(( $(date +%s) - s ))"; grep -v " OK" gpurun_out/ref_$t.log | head -6
done
