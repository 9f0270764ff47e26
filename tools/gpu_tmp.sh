#!/bin/bash
mkdir -p gpurun_out
timeout 150 python tools/variant_check.py 0,4 1073741824 0 > gpurun_out/r2_variant_check7.log 2>&1; echo "variant rc=$?"; tail -n 6 gpurun_out/r2_variant_check7.log
timeout 100 python tools/dbg_timing.py 8,9 > gpurun_out/r2_dbg_timing5.log 2>&1; echo "rc=$?"; cat gpurun_out/r2_dbg_timing5.log
