#!/bin/bash
# Runs the reference's own test binaries (built by simdutf_b200.build.build_reference_integration) against
# the b200 implementation.  usage: tools/run_ref_tests.sh [outdir]
D=simdutf_b200/build/with_b200
OUT=${1:-gpurun_out/ref_tests}
mkdir -p "$OUT"
for t in $(ls $D | grep -vE '\.(o|a|so|inc|cpp)$'); do
  start=$(date +%s.%N)
  timeout 600 $D/$t -a b200 > "$OUT/$t.log" 2>&1
  rc=$?
  end=$(date +%s.%N)
  printf "%-50s rc=%d %s\n" "$t" "$rc" "$(grep -c OK "$OUT/$t.log") OK lines; last: $(tail -n 1 "$OUT/$t.log" | cut -c1-80)"
done
