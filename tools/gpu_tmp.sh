#!/bin/bash
mkdir -p gpurun_out
timeout 400 python tools/prof_sleep.py 1073741824 5 > gpurun_out/r2_scan_sleep.log 2>&1; echo "rc=$?"; tail -n 12 gpurun_out/r2_scan_sleep.log
