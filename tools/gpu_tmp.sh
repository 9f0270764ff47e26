#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "golden or utf8_small or utf8_error or utf8_medium or bitplane" > gpurun_out/r2_len_parity.log 2>&1; echo "parity rc=$?"; tail -n 3 gpurun_out/r2_len_parity.log
timeout 100 python tools/prof_one.py length 1073741824 10 2>&1 | tail -1
