#!/bin/bash
# compute-sanitizer memcheck + racecheck of the hot path on a 3072-atom cell; summaries go to gpurun_out/ (copied to profiles/)
OUT=gpurun_out; mkdir -p $OUT
for tool in memcheck racecheck; do
  timeout 1500 compute-sanitizer --tool $tool --print-limit 20 --log-file $OUT/sanitizer_$tool.log python tests/gpu_sanitize_target.py > $OUT/sanitizer_$tool.out 2>&1
  echo "== $tool rc=$?"; tail -3 $OUT/sanitizer_$tool.out; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|hazard" $OUT/sanitizer_$tool.log | tail -5
done
