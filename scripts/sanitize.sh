#!/bin/bash
# compute-sanitizer over the cross-CTA kernels at small shapes (race / sync / memory checking).
# usage (under gpurun): scripts/sanitize.sh <tag>      -> gpurun_out/<tag>_sanitizer_<tool>_<target>.log
set -u
TAG=${1:-r02}
OUT=gpurun_out
mkdir -p $OUT
CS=/usr/local/cuda/bin/compute-sanitizer
for tool in memcheck racecheck synccheck; do
  for target in lstm gemm x3 step; do
    log=$OUT/${TAG}_sanitizer_${tool}_${target}.log
    timeout 420 $CS --tool $tool --print-limit 20 --error-exitcode 9 python scripts/sanitize_target.py $target > $log 2>&1
    rc=$?
    echo "$tool $target rc=$rc $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' $log | tail -1)" | tee -a $OUT/${TAG}_sanitizer_summary.log
  done
done
