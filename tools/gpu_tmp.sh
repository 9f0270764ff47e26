#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_reference_suite.py -m gpu -x -q -k "utf32_family or latin1_family or golden or convert_utf32_to_utf8 or convert_utf32_to_utf16 or convert_utf16le_to_utf32 or convert_utf16be_to_utf32 or convert_latin1_to_utf8 or convert_utf8_to_latin1 or basic_fuzzer or special_tests" --durations=5 > gpurun_out/r2_k8v3_full.log 2>&1; echo "rc=$?"; tail -n 12 gpurun_out/r2_k8v3_full.log
for op in utf32to8 utf32to16 utf16to32 l1to8 u8tol1; do
timeout 100 python tools/prof_one.py $op 1073741824 5 2>&1 | tail -1
done
