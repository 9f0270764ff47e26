#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "batch" > gpurun_out/r2_batch.log 2>&1; echo "batch rc=$?"; tail -n 15 gpurun_out/r2_batch.log
timeout 100 python tools/prof_one.py validate_mixed 1073741824 10 > gpurun_out/r2_k1_mixed.log 2>&1; tail -n 1 gpurun_out/r2_k1_mixed.log
timeout 100 python tools/prof_one.py validate_ascii 1073741824 10 > gpurun_out/r2_k1_ascii.log 2>&1; tail -n 1 gpurun_out/r2_k1_ascii.log
timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "well_formed or utf16be or utf16_random" > gpurun_out/r2_wf.log 2>&1; echo "wf rc=$?"; tail -n 3 gpurun_out/r2_wf.log
