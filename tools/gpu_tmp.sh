#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_full_size_vs_reference.py tests/test_reference_suite.py -m gpu -x -q -k "golden or utf16_random or bitplane or utf16be_twins or config3 or utf16le_to_utf8 or utf16be_to_utf8 or utf16 and not to_well" > gpurun_out/r2_k6v3_full.log 2>&1; echo "rc=$?"; tail -n 6 gpurun_out/r2_k6v3_full.log
timeout 100 python tools/prof_one.py utf16to8 2147483648 5 2>&1 | tail -1
