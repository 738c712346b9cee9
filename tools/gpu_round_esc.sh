#!/bin/bash
# (1) parity of the CTA backward pass after the single-code-path change; (2) per-phase tick log of car_escape at B=2048; (3) tail timing
TAG=${1:-r01e2}
OUT=gpurun_out
mkdir -p $OUT
TRAJOPT_B200_BP_CTA_THRESHOLD=100000000 timeout 600 python tools/gpu_diag.py > $OUT/${TAG}_diag_cta_forced.log 2>&1; tail -1 $OUT/${TAG}_diag_cta_forced.log
TRAJOPT_B200_TICK_DETAIL=1 TRAJOPT_B200_TICK_LOG=$OUT/${TAG}_ticks_quad.txt timeout 300 python bench.py --batch 16384 --steps 1 --warmup 0 --no-cpu-baseline > $OUT/${TAG}_bench_quad.json 2> $OUT/${TAG}_bench_quad.err
grep 'timed step' $OUT/${TAG}_bench_quad.err; python tools/tick_summary.py $OUT/${TAG}_ticks_quad.txt
TRAJOPT_B200_TICK_DETAIL=1 TRAJOPT_B200_TICK_LOG=$OUT/${TAG}_ticks_escape.txt timeout 600 python tools/run_configs.py ${TAG} 8 escape_altro > $OUT/${TAG}_escape.log 2>&1
python tools/tick_summary.py $OUT/${TAG}_ticks_escape.txt
TRAJOPT_B200_TICK_DETAIL=1 TRAJOPT_B200_TICK_LOG=$OUT/${TAG}_ticks_park.txt timeout 600 python tools/run_configs.py ${TAG}p 8 park_inf_altro > $OUT/${TAG}_park.log 2>&1
python tools/tick_summary.py $OUT/${TAG}_ticks_park.txt
