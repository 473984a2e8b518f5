#!/bin/sh
# compute-sanitizer over tools/sanitize_targets.py; logs under gpurun_out/ (copy the summaries to profiles/).
#   tools/run_sanitizers.sh [memcheck racecheck synccheck initcheck]
cd "$(dirname "$0")/.."
TOOLS=${*:-"memcheck racecheck"}
mkdir -p gpurun_out
for t in $TOOLS; do
  for part in bbduk direct steps kcount; do
    log=gpurun_out/r02_sanitizer_${t}_${part}.txt
    SAN_PAIRS=${SAN_PAIRS:-600} timeout 1500 compute-sanitizer --tool $t --print-limit 20 --error-exitcode 3 \
        python tools/sanitize_targets.py $part > $log 2>&1
    echo "$t $part rc=$? $(grep -c 'ok' $log) ok-lines; $(grep 'ERROR SUMMARY\|RACECHECK SUMMARY' $log | tail -1)"
  done
done
