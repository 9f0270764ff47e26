#!/bin/bash
mkdir -p gpurun_out
for v in 4 5 3 0 6 1 2; do
echo "variant $v"
B200_BENCH_TUNE="conv_minb=$v" timeout 100 python tools/prof_one.py validate_mixed 1073741824 10 2>&1 | tail -1
B200_BENCH_TUNE="conv_minb=$v" timeout 100 python tools/prof_one.py validate_ascii 1073741824 10 2>&1 | tail -1
done
timeout 100 python tools/prof_one.py wellformed 1073741824 5 2>&1 | tail -1
timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "well_formed or utf8_small_random or utf8_error_classes or utf8_medium or golden" 2>&1 | tail -2
