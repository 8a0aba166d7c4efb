#!/bin/bash
mkdir -p gpurun_out
for tool in memcheck racecheck; do
  timeout 1200 compute-sanitizer --tool $tool --kernel-name-exclude kns=at --print-limit 20 python tools/sanitize_train.py > gpurun_out/sanitize_train_$tool.log 2>&1
  echo "== $tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitize_train ok|Invalid|hazard" gpurun_out/sanitize_train_$tool.log | sort | uniq -c | head -12
done
