timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "utf8 or golden or config2 or repeated or edge" 2>&1 | tail -3
python tools/prof_texts.py 2>&1 | head -2
python tools/prof_one.py convert16 1073741824 5 2>&1 | tail -n 1
