python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -5
for mb in 2 3 4; do echo "MINB=$mb $(B200_TUNE_MINB=$mb python tools/prof_one.py utf16to8 2147483648 5 2>&1 | tail -n 1)"; done
