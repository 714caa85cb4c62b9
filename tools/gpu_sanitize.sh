#!/bin/bash
# compute-sanitizer over tools/sanitize_case.py; summaries into gpurun_out/<tag>_sanitizer_*.txt
TAG=${1:-r01}
mkdir -p gpurun_out
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_case.py > gpurun_out/${TAG}_sanitizer_$tool.log 2>&1
  echo "exit $?" >> gpurun_out/${TAG}_sanitizer_$tool.log
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitize case ok|exit|Error|error|hazard" gpurun_out/${TAG}_sanitizer_$tool.log | sort | uniq -c | head -20 > gpurun_out/${TAG}_sanitizer_$tool.txt
  cat gpurun_out/${TAG}_sanitizer_$tool.txt
done
