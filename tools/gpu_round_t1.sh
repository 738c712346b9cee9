#!/bin/bash
TAG=${1:-r01t1}
OUT=gpurun_out
mkdir -p $OUT
run() {
  name=$1; bb=$2; shift; shift
  env "$@" TRAJOPT_B200_TICK_DETAIL=1 TRAJOPT_B200_TICK_LOG=$OUT/${TAG}_ticks_$name.txt timeout 300 python bench.py --batch $bb --steps 1 --warmup 0 --no-cpu-baseline > $OUT/${TAG}_bench_$name.json 2> $OUT/${TAG}_bench_$name.err
  echo "== $name: $(grep 'timed step' $OUT/${TAG}_bench_$name.err)"
  python tools/tick_summary.py $OUT/${TAG}_ticks_$name.txt | tail -3
}
run base 4096 A=1
run all1 4096 TRAJOPT_B200_TRIAL_ALL_MINB=1
TRAJOPT_B200_TRIAL_ALL_MINB=1 timeout 600 python tools/gpu_diag.py quad_altro quad_regdiv > $OUT/${TAG}_diag_all1.log 2>&1; tail -1 $OUT/${TAG}_diag_all1.log
timeout 900 python tools/gpu_diag.py escape_notebook park_inf_altro car_3obs_altro > $OUT/${TAG}_diag_rows.log 2>&1; tail -1 $OUT/${TAG}_diag_rows.log
