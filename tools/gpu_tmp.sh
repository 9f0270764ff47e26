ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"^k_" -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-extras --e2e-steps 1 --cpu-sample-bytes 33554432 > gpurun_out/ncu_bench.log 2>&1; echo "ncu rc=$?"
wc -l gpurun_out/launches.csv; tail -3 gpurun_out/ncu_bench.log
