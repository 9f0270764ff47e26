#!/bin/bash
# Full GPU check: every -m gpu test (parity, full-size configs against the reference, the reference's own binaries,
# the sharded path), smoke, then bench (own arm + reference arm).
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -m gpu -q --durations=25 > gpurun_out/r02_pytest_gpu_full.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r02_pytest_gpu_full.log
tail -n 8 gpurun_out/r02_pytest_gpu_full.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.log 2>&1; echo "smoke rc=$?"; tail -n 2 gpurun_out/r02_smoke.log
timeout 600 python bench.py > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err; echo "bench rc=$?"; cut -c1-400 gpurun_out/r02_bench.json; tail -n 3 gpurun_out/r02_bench.err
timeout 300 python bench.py --impl reference > gpurun_out/r02_bench_ref.json 2> gpurun_out/r02_bench_ref.err; echo "bench ref rc=$?"; cut -c1-300 gpurun_out/r02_bench_ref.json
