#!/bin/bash
mkdir -p gpurun_out
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29527 bench.py --gpus 4 --steps 20 --warmup 3 > gpurun_out/r02_bench_4gpu.json 2> gpurun_out/r02_bench_4gpu.err; echo "bench4 rc=$?"; cut -c1-300 gpurun_out/r02_bench_4gpu.json; tail -n 3 gpurun_out/r02_bench_4gpu.err
