timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -5
echo "fused:   $(timeout 120 python tools/prof_one.py convert16 1073741824 5 2>&1 | tail -n 1)"
echo "unfused: $(B200_TUNE_FUSED=0 timeout 120 python tools/prof_one.py convert16 1073741824 5 2>&1 | tail -n 1)"
for lead in 64 128 512 1024; do echo "lead=$lead: $(B200_TUNE_LEAD=$lead timeout 120 python tools/prof_one.py convert16 1073741824 5 2>&1 | tail -n 1)"; done
echo "fused32: $(timeout 120 python tools/prof_one.py convert32 1073741824 5 2>&1 | tail -n 1)"
