#!/bin/bash
# compute-sanitizer over every kernel family.  One GPU, under gpurun:
#   bash scripts/sanitize_r02.sh   -> gpurun_out/r02_sanitize_{memcheck,racecheck,synccheck,initcheck}.log
mkdir -p gpurun_out
for tool in ${SANITIZE_TOOLS:-memcheck racecheck synccheck initcheck}; do
  timeout ${SANITIZE_TIMEOUT:-420} compute-sanitizer --tool $tool --print-limit 20 --log-file gpurun_out/r02_sanitize_$tool.log \
      python scripts/sanitize_ops.py > gpurun_out/r02_sanitize_$tool.out 2>&1
  echo "== $tool: exit $? =="; tail -2 gpurun_out/r02_sanitize_$tool.out; grep -E "ERROR SUMMARY|RACECHECK SUMMARY" gpurun_out/r02_sanitize_$tool.log | tail -2
done
