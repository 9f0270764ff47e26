#!/bin/bash
# usage: gpu_ncu_one.sh <op> <kernel-regex> <tag> [bytes]   — plain run first, then ONE ncu --set full capture of the kernel
op=$1; kre=$2; tag=$3; bytes=${4:-268435456}
mkdir -p gpurun_out
python tools/prof_one.py $op $bytes 3 > gpurun_out/plain_$tag.log 2>&1 || { tail gpurun_out/plain_$tag.log; exit 1; }
tail -n 1 gpurun_out/plain_$tag.log
ncu --set full --clock-control none --import-source on -k regex:$kre -s 2 -c 1 -f -o gpurun_out/prof_$tag \
  python tools/prof_one.py $op $bytes 3 > gpurun_out/ncu_$tag.log 2>&1
echo "ncu rc=$?"; ls -la gpurun_out/prof_$tag.ncu-rep
