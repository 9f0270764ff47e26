timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "latin1_family or utf16 or well_formed or golden" 2>&1 | tail -3
timeout 200 python tools/detect_probe.py 2>&1 | tail -4
for op in l1to8 u8tol1; do timeout 200 python tools/prof_one.py $op $((1<<29)) 5 2>&1 | tail -1; done
