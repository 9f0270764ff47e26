#!/bin/bash
mkdir -p gpurun_out
timeout 700 python -m pytest tests/test_gpu_parity.py tests/test_reference_suite.py -m gpu -x -q -k "base64 or utf32_family or well_formed_and_detect or golden or detect_encodings or validate_utf32_with_errors" --durations=8 > gpurun_out/r2_b64v3_tests.log 2>&1; echo "pytest rc=$?"; tail -n 25 gpurun_out/r2_b64v3_tests.log
timeout 500 python tools/prof_b64.py 1073741824 5 > gpurun_out/r2_b64v3_prof.log 2>&1; echo "prof rc=$?"; cat gpurun_out/r2_b64v3_prof.log | tail -n 60
