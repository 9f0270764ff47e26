timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "latin1_family" 2>&1 | tail -3
for op in l1to8 u8tol1; do timeout 200 python tools/prof_one.py $op $((1<<29)) 5 2>&1 | tail -1; done
timeout 300 simdutf_b200/build/with_b200/convert_latin1_to_utf8_tests -a b200 2>&1 | tail -2
timeout 300 simdutf_b200/build/with_b200/convert_utf8_to_latin1_tests -a b200 2>&1 | tail -2
