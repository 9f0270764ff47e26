timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "base64" 2>&1 | tail -12
timeout 900 simdutf_b200/build/with_b200/base64_tests -a b200 > gpurun_out/ref_base64_tests.log 2>&1; echo "base64_tests rc=$? OK=$(grep -c ' OK' gpurun_out/ref_base64_tests.log)"; grep -v " OK" gpurun_out/ref_base64_tests.log | head -12
