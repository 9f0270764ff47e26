timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "edge_paths or base64 or utf16_random" 2>&1 | tail -25
