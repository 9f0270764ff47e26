timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "utf32_family" 2>&1 | tail -3
for op in utf32to8 utf32to16 utf32to16be; do timeout 200 python tools/prof_one.py $op $((1<<30)) 5 2>&1 | tail -1; done
