#!/bin/bash
# scratch: first run of the single-pass transcoder + the new host layer
mkdir -p gpurun_out
nvidia-smi -L
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "golden or utf8_small or utf8_error or utf8_medium or bitplane or utf16be or utf32_family or repeated or config2 or beyond_4gib or host_streaming" > gpurun_out/r2_k3sp_parity.log 2>&1; echo "parity rc=$?"; tail -n 15 gpurun_out/r2_k3sp_parity.log
timeout 900 python -m pytest tests/test_sharded_gpu.py -m gpu -x -q > gpurun_out/r2_sharded_test2.log 2>&1; echo "sharded rc=$?"; tail -n 15 gpurun_out/r2_sharded_test2.log
for mb in 3 4; do
  B200_BENCH_TUNE=conv_minb=$mb timeout 600 python bench.py --steps 20 --warmup 3 --no-extras > gpurun_out/r2_bench_sp_minb$mb.json 2> gpurun_out/r2_bench_sp_minb$mb.err; echo "bench minb=$mb rc=$?"; python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2_bench_sp_minb$mb.json"))
    print({k:d[k] for k in ("value","ms_per_step")}, {k:d["roofline"][k] for k in ("frac","avg_launch_ms","length_kernel_ms","validate_utf8_ascii_frac","validate_utf8_mixed_frac")}, d["e2e"]["value"], d["cpu_baseline"]["value"] if d["cpu_baseline"] else None)
except Exception as e:
    print("no json", e)
PY
  tail -n 3 gpurun_out/r2_bench_sp_minb$mb.err
done
