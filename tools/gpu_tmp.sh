mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q --durations=40 > gpurun_out/pytest_gpu_full.log 2>&1; echo "pytest rc=$?"
tail -n 60 gpurun_out/pytest_gpu_full.log
