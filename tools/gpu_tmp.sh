#!/bin/bash
# scratch: warp-specialised single-pass transcoder, mbarrier hand-offs
mkdir -p gpurun_out
timeout 180 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1; rc=$?; echo "smoke rc=$rc"; tail -n 3 gpurun_out/r2_smoke.log
if [ $rc -ne 0 ]; then exit 1; fi
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "golden or utf8_small or utf8_error or utf8_medium or bitplane or utf16be or utf32_family or repeated or config2 or beyond_4gib or host_streaming" > gpurun_out/r2_k3sp_parity.log 2>&1; echo "parity rc=$?"; tail -n 5 gpurun_out/r2_k3sp_parity.log
for mb in 3 4; do
  B200_BENCH_TUNE=conv_minb=$mb python tools/prof_one.py convert16 1073741824 10 2>&1 | tail -n 1
done
export B200_BENCH_TUNE=conv_minb=4
ncu --set full --clock-control none --import-source on -k regex:k_utf8_transcode_sp -s 2 -c 1 -f -o gpurun_out/r2_sp_v2_minb4 python tools/prof_one.py convert16 268435456 3 > gpurun_out/ncu_sp_minb4.log 2>&1
echo "ncu rc=$?"
