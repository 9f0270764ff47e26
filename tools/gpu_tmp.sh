timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "latin1_family or utf32_family" 2>&1 | tail -25
for op in l1to8 l1to16 l1to32 u8tol1 u16tol1 u32tol1 validate_ascii_op len8froml1; do timeout 300 python tools/prof_one.py $op $((1<<29)) 5 2>&1 | tail -2; done
timeout 1500 python -m pytest tests/test_reference_suite.py -m gpu -q -k "latin1 or ascii or bele" 2>&1 | tail -15
timeout 600 simdutf_b200/build/with_b200/random_fuzzer -a b200 > gpurun_out/ref_random_fuzzer.log 2>&1; echo "random_fuzzer rc=$?"; tail -5 gpurun_out/ref_random_fuzzer.log
