#!/bin/bash
# source-level ncu captures: one bulk tick (8,192 live problems) and one tail tick (64-problem batch, late tick)
TAG=${1:-r01p}
OUT=gpurun_out
mkdir -p $OUT
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'ls_' -s 60 -c 8 -f -o $OUT/${TAG}_prof_bulk \
    python bench.py --batch 8192 --steps 1 --warmup 0 --no-cpu-baseline > $OUT/${TAG}_ncu_bulk.log 2>&1
echo "ncu bulk exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'ls_' -s 700 -c 7 -f -o $OUT/${TAG}_prof_tail \
    python bench.py --batch 64 --steps 1 --warmup 0 --no-cpu-baseline > $OUT/${TAG}_ncu_tail.log 2>&1
echo "ncu tail exit $?"
