#!/bin/bash
# compute-sanitizer over the kernels with hand-rolled synchronisation (the
# reference has no sanitizer targets; here the mbarrier / TMEM pipelines and the two-stream executor
# are exactly where memcheck / racecheck / synccheck pay off).  Run on a GPU box:
#   gpurun --timeout 1500 -- 'bash scripts/sanitize.sh gpurun_out/sanitize'
# Writes one log per tool and a summary line per tool to $OUT/summary.txt; exit code 0 only if every
# tool reports 0 errors.
OUT=${1:-gpurun_out/sanitize}
mkdir -p "$OUT"
rc_all=0
for tool in memcheck racecheck synccheck; do
  SANITIZE_N=${SANITIZE_N:-1500} timeout ${SANITIZE_TIMEOUT:-420} compute-sanitizer --tool $tool \
      --error-exitcode 3 --print-limit 20 python scripts/sanitize_cases.py > "$OUT/$tool.log" 2>&1
  rc=$?
  echo "$tool: rc=$rc $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' "$OUT/$tool.log" | tail -1)" | tee -a "$OUT/summary.txt"
  [ $rc -ne 0 ] && rc_all=1
done
exit $rc_all
