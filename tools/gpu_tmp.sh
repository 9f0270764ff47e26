python tools/e2e_probe.py raw 2>&1 | tail -3
for s in 8 16 64 128; do B200_TUNE_SEG_MB=$s python tools/e2e_probe.py 2>&1 | tail -1; done
