#!/bin/bash
# compute-sanitizer over the small-shape target (memcheck, then racecheck); summaries -> gpurun_out/
mkdir -p gpurun_out
for TOOL in memcheck racecheck; do
  timeout ${SANITIZE_TIMEOUT:-420} compute-sanitizer --tool $TOOL --print-limit 20 python tools/sanitize_target.py \
      > gpurun_out/sanitize_$TOOL.log 2>&1
  echo "== $TOOL rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitize target ok|Error|hazard" gpurun_out/sanitize_$TOOL.log | head -12
done
