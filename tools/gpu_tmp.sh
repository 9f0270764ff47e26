timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -15
timeout 120 python tools/prof_one.py base64 2147483648 5 2>&1 | tail -n 1
