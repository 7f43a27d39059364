#!/bin/bash
# tools/sanitize.sh <tag>: compute-sanitizer memcheck + initcheck (+ racecheck on shared memory) of tools/sanitizer_smoke.py; summaries -> gpurun_out/<tag>_sanitizer_*.log
tag=$1
for tool in memcheck initcheck racecheck; do
  n=200; [ $tool = racecheck ] && n=40
  timeout 1200 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitizer_smoke.py 2048 $n > gpurun_out/${tag}_sanitizer_${tool}.log 2>&1
  echo "== $tool: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/${tag}_sanitizer_${tool}.log | tail -1)  $(grep -c 'sanitizer smoke ok' gpurun_out/${tag}_sanitizer_${tool}.log) ok-lines"
done
