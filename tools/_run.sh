python tools/sanitizer_smoke.py > gpurun_out/san_plain.log 2>&1 || { tail -5 gpurun_out/san_plain.log; exit 1; }
for tool in memcheck racecheck; do
  timeout 600 compute-sanitizer --tool $tool --error-exitcode 7 python tools/sanitizer_smoke.py > gpurun_out/san_$tool.log 2>&1; echo "$tool exit $?"
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitizer smoke done|=========   " gpurun_out/san_$tool.log | tail -5
done
