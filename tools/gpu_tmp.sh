for mb in 3; do echo "K=2 MINB=$mb: $(B200_TUNE_K=2 B200_TUNE_MINB=$mb python tools/prof_one.py convert16 1073741824 5 2>&1 | tail -n 1)"; done
python tools/prof_one.py convert32 1073741824 5 2>&1 | tail -n 1
