#!/bin/bash
TAG=${1:-r01s}
TOOL=${2:-racecheck}
OUT=gpurun_out
mkdir -p $OUT
timeout 1500 compute-sanitizer --tool $TOOL --print-limit 20 python tools/sanitize.py di_altro pend_mintime park_inf_altro pend_sqrt_altro > $OUT/${TAG}_${TOOL}.log 2>&1; echo "sanitizer exit $?"
grep -E "ERROR SUMMARY|RACECHECK SUMMARY|hazard|Invalid" $OUT/${TAG}_${TOOL}.log | head -20
tail -5 $OUT/${TAG}_${TOOL}.log
