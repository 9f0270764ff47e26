#!/bin/bash
# GPU-side sweep of the UTF-8 -> UTF-16 kernel variants: parity tests first, then timings.
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "utf8 or golden or config2 or repeated" > gpurun_out/pytest_conv.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_conv.log
tail -n 15 gpurun_out/pytest_conv.log
for k in 1 2 4; do for mb in 1 2 3; do [ "$k$mb" = "43" -o "$k$mb" = "11" ] && continue
  echo "K=$k MINB=$mb: $(B200_TUNE_K=$k B200_TUNE_MINB=$mb python tools/prof_one.py convert16 1073741824 5 2>&1 | tail -n 1)"
done; done | tee gpurun_out/conv_sweep.log
