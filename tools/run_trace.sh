#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 150 python tests/soak.py 60 2>&1 | tail -2 | tee gpurun_out/r2_soak.txt
for tool in memcheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool python tests/sanitizer_run.py > gpurun_out/r2_sanitizer_$tool.txt 2>&1; echo "$tool rc=$?"; grep -E "ERROR SUMMARY|lite|both|resident|single|192" gpurun_out/r2_sanitizer_$tool.txt | tail -12
done
