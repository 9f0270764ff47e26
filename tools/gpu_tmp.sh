#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "golden or utf8_small or utf8_error or utf8_medium or bitplane or utf16be or utf32_family or repeated or beyond_4gib or host_streaming or config2 or single_pass" > gpurun_out/r2_k3_k3_parity.log 2>&1; echo "parity rc=$?"; tail -n 3 gpurun_out/r2_k3_k3_parity.log
timeout 100 python tools/prof_one.py convert16 1073741824 10 2>&1 | tail -1
timeout 100 python tools/prof_one.py convert32 1073741824 10 2>&1 | tail -1
timeout 100 python tools/dbg_timing.py 8 > gpurun_out/r2_dbg_timing_final.log 2>&1; tail -n 4 gpurun_out/r2_dbg_timing_final.log
timeout 200 python tools/prof_texts.py > gpurun_out/r2_text_shapes.log 2>&1; tail -n 7 gpurun_out/r2_text_shapes.log
