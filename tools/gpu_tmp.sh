#!/bin/bash
mkdir -p gpurun_out
timeout 500 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "single_pass_transcoder_structure or batch" --durations=5 > gpurun_out/r2_k3_struct.log 2>&1; echo "rc=$?"; tail -n 12 gpurun_out/r2_k3_struct.log
