#!/bin/bash
TAG=${1:-r01k2}
OUT=gpurun_out
mkdir -p $OUT
run() {
  name=$1; bb=$2; shift; shift
  env "$@" TRAJOPT_B200_TICK_DETAIL=1 TRAJOPT_B200_TICK_LOG=$OUT/${TAG}_ticks_$name.txt timeout 300 python bench.py --batch $bb --steps 1 --warmup 0 --no-cpu-baseline > $OUT/${TAG}_bench_$name.json 2> $OUT/${TAG}_bench_$name.err
  echo "== $name: $(grep 'timed step' $OUT/${TAG}_bench_$name.err)"
  python tools/tick_summary.py $OUT/${TAG}_ticks_$name.txt | head -3
}
run k0 65536 TRAJOPT_B200_BP_INLINE_RESTARTS=0
run k1 65536 TRAJOPT_B200_BP_INLINE_RESTARTS=1
run k2 65536 TRAJOPT_B200_BP_INLINE_RESTARTS=2
run k4 65536 TRAJOPT_B200_BP_INLINE_RESTARTS=4
TRAJOPT_B200_BP_INLINE_RESTARTS=0 TRAJOPT_B200_TAIL_THRESHOLD=0 TRAJOPT_B200_BP_CTA_THRESHOLD=0 timeout 900 python tools/gpu_diag.py quad_altro quad_regdiv quad_obs_al cart_altro > $OUT/${TAG}_diag_k0.log 2>&1; tail -1 $OUT/${TAG}_diag_k0.log
