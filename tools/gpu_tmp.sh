#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "base64 or golden or utf8_small_random or utf8_error_classes or well_formed_and_detect" > gpurun_out/r2_b64v3_tests2.log 2>&1; echo "pytest rc=$?"; tail -n 4 gpurun_out/r2_b64v3_tests2.log
timeout 200 python tools/prof_b64.py 1073741824 5 0 > gpurun_out/r2_b64v3_prof2.log 2>&1; echo "prof rc=$?"; grep -A1 -E "base64|validate_utf8 on|detect" gpurun_out/r2_b64v3_prof2.log
bash tools/gpu_ncu_one.sh base64 k_b64_decode_v3 r02_k7_b64_decode_1GiB 1073741824
bash tools/gpu_ncu_one.sh utf16to8 k_utf16_to_utf8_v3 r02_k6_utf16_to_utf8_2GiB 2147483648
