#!/bin/bash
# CTA backward pass: occupancy variants x threshold, group-kernel MINB=4, and a 65,536 tick log
TAG=${1:-r01r}
OUT=gpurun_out
mkdir -p $OUT
run() {
  name=$1; bb=$2; shift; shift
  env "$@" TRAJOPT_B200_TICK_DETAIL=1 TRAJOPT_B200_TICK_LOG=$OUT/${TAG}_ticks_$name.txt timeout 300 python bench.py --batch $bb --steps 1 --warmup 0 --no-cpu-baseline > $OUT/${TAG}_bench_$name.json 2> $OUT/${TAG}_bench_$name.err
  echo "== $name: $(grep 'timed step' $OUT/${TAG}_bench_$name.err)"
  python tools/tick_summary.py $OUT/${TAG}_ticks_$name.txt
}
run m2_t4096 16384 TRAJOPT_B200_BP_CTA_THRESHOLD=4096
run m3_t4096 16384 TRAJOPT_B200_BP_CTA_MINB=3 TRAJOPT_B200_BP_CTA_THRESHOLD=4096
run m4_t4096 16384 TRAJOPT_B200_BP_CTA_MINB=4 TRAJOPT_B200_BP_CTA_THRESHOLD=4096
run m3_all 16384 TRAJOPT_B200_BP_CTA_MINB=3 TRAJOPT_B200_BP_CTA_THRESHOLD=100000000
run m4_all 16384 TRAJOPT_B200_BP_CTA_MINB=4 TRAJOPT_B200_BP_CTA_THRESHOLD=100000000
run bpminb4 16384 TRAJOPT_B200_BP_MINB=4 TRAJOPT_B200_BP_CTA_THRESHOLD=4096
run big 65536 TRAJOPT_B200_BP_CTA_THRESHOLD=4096
TRAJOPT_B200_BP_CTA_MINB=4 TRAJOPT_B200_BP_CTA_THRESHOLD=100000000 timeout 600 python tools/gpu_diag.py quad_altro quad_ilqr quad_regdiv > $OUT/${TAG}_diag_m4.log 2>&1; tail -1 $OUT/${TAG}_diag_m4.log
