timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "utf16 or golden or well_formed or config3 or bitplane" 2>&1 | tail -5
timeout 200 python tools/detect_probe.py 2>&1 | tail -5
for t in validate_utf16le_basic_tests validate_utf16le_with_errors_tests validate_utf16be_with_errors_tests; do timeout 200 simdutf_b200/build/with_b200/$t -a b200 > gpurun_out/ref_$t.log 2>&1; echo "$t rc=$? OK=$(grep -c ' OK' gpurun_out/ref_$t.log)"; done
