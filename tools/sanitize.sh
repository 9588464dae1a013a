#!/bin/bash
# compute-sanitizer over the reduced parity subset (tools/sanitize_cases.py).  ONE tool per gpurun call
# (B200_PROFILING.md): gpurun --timeout 1500 -- tools/sanitize.sh memcheck|racecheck|synccheck|initcheck
set -u
TOOL=${1:-memcheck}
OUT=gpurun_out/sanitize; mkdir -p $OUT
timeout 120 python tools/sanitize_cases.py > $OUT/plain.log 2>&1 || { echo "plain run failed"; tail -5 $OUT/plain.log; exit 1; }
EXTRA=""
[ "$TOOL" = "racecheck" ] && EXTRA="--racecheck-report analysis"
timeout 1300 compute-sanitizer --tool $TOOL $EXTRA --print-limit 30 --log-file $OUT/$TOOL.log python tools/sanitize_cases.py > $OUT/${TOOL}_stdout.log 2>&1
echo "rc=$?"
tail -3 $OUT/${TOOL}_stdout.log
grep -c "=========" $OUT/$TOOL.log; tail -25 $OUT/$TOOL.log
