python tools/sanitize_smoke.py 2>&1 | tail -2
for tool in memcheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_smoke.py > gpurun_out/sanitizer_$tool.log 2>&1; echo "$tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|Error|error" gpurun_out/sanitizer_$tool.log | head -8
done
