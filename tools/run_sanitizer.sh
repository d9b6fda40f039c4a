#!/bin/bash
# compute-sanitizer memcheck + racecheck of the small end-to-end script; logs under gpurun_out/
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for tool in memcheck racecheck; do
  timeout 1500 compute-sanitizer --tool $tool --target-processes all --print-limit 20 python tools/sanitize_small.py > gpurun_out/r02_sanitizer_$tool.log 2>&1
  echo "$tool rc=$?" >> gpurun_out/r02_sanitizer_$tool.log
  tail -6 gpurun_out/r02_sanitizer_$tool.log
done
