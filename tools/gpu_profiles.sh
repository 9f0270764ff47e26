#!/bin/bash
# Round-2 evidence: one ncu --set full capture per headline kernel (1 GiB inputs, after a plain run of the same command),
# and the ncu launch list of a short bench.py run.  Outputs under gpurun_out/; tools/ncu_summary.py turns the reports
# into profiles/r02_*.json.
mkdir -p gpurun_out
cap() {  # op kernel-regex tag bytes
  timeout 120 python tools/prof_one.py $1 $4 5 > gpurun_out/plain_$3.log 2>&1 || { tail -n 3 gpurun_out/plain_$3.log; return; }
  tail -n 1 gpurun_out/plain_$3.log
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$2 -s 2 -c 1 -f -o gpurun_out/$3 \
    python tools/prof_one.py $1 $4 3 > gpurun_out/ncu_$3.log 2>&1
  echo "ncu $3 rc=$?"
}
cap convert16 k_utf8_transcode_v3 r02_k3_convert16_1GiB 1073741824
cap validate_ascii k_validate_utf8 r02_k1_validate_ascii_1GiB 1073741824
cap validate_mixed k_validate_utf8 r02_k1_validate_mixed_1GiB 1073741824
cap length k_count_utf8 r02_k2_utf16_length_1GiB 1073741824
cap utf16to8 k_utf16_to_utf8_v3 r02_k6_utf16_to_utf8_2GiB 2147483648
timeout 300 python bench.py --steps 2 --warmup 3 --no-extras > gpurun_out/r02_bench_short.json 2> gpurun_out/r02_bench_short.err; echo "bench short rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:::k_ -c 400 --csv --log-file gpurun_out/r02_bench_launch_list.csv \
  python bench.py --steps 2 --warmup 3 --no-extras > gpurun_out/r02_bench_under_ncu.log 2>&1; echo "launch list rc=$?"
