#!/bin/bash
# quick check of a kernel change: parity diag (default + grouped line search), quad tick timing at 16,384
TAG=${1:-r01t}
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python tools/gpu_diag.py > $OUT/${TAG}_diag.log 2>&1; tail -1 $OUT/${TAG}_diag.log
TRAJOPT_B200_TAIL_THRESHOLD=0 TRAJOPT_B200_BP_CTA_THRESHOLD=0 timeout 600 python tools/gpu_diag.py quad_altro quad_regdiv cart_altro escape_notebook park_inf_altro pend_mintime > $OUT/${TAG}_diag_grouped.log 2>&1; tail -1 $OUT/${TAG}_diag_grouped.log
TRAJOPT_B200_TICK_DETAIL=1 TRAJOPT_B200_TICK_LOG=$OUT/${TAG}_ticks_quad.txt timeout 300 python bench.py --batch 16384 --steps 1 --warmup 0 --no-cpu-baseline > $OUT/${TAG}_bench_quad.json 2> $OUT/${TAG}_bench_quad.err
grep 'timed step' $OUT/${TAG}_bench_quad.err; python tools/tick_summary.py $OUT/${TAG}_ticks_quad.txt
