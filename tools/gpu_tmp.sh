timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
python tools/prof_one.py convert16 1073741824 5 2>&1 | tail -n 1
python tools/prof_one.py convert32 1073741824 5 2>&1 | tail -n 1
