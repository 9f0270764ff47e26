python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
python tools/prof_one.py validate_mixed 1073741824 5 2>&1 | tail -n 1
python tools/prof_one.py validate_ascii 1073741824 5 2>&1 | tail -n 1
echo "K=2 MINB=4: $(B200_TUNE_K=2 B200_TUNE_MINB=4 python tools/prof_one.py convert16 1073741824 5 2>&1 | tail -n 1)"
