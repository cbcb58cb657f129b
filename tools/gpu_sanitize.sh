#!/bin/bash
set -u
mkdir -p gpurun_out
for tool in memcheck synccheck racecheck; do
  timeout 420 compute-sanitizer --tool $tool --print-limit 20 python tests/dev/sanitize_run.py > gpurun_out/sanitize_$tool.log 2>&1; echo "$tool rc=$?"
  grep -E "ERROR SUMMARY|sanitize_run ok|RACECHECK SUMMARY|Error|hazard" gpurun_out/sanitize_$tool.log | head -12
done
