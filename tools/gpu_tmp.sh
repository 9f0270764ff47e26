timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "latin1_family" 2>&1 | tail -12
