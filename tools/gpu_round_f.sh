#!/bin/bash
# kernel-variant sweep with per-phase tick timing
TAG=${1:-r01f}
BB=${2:-16384}
OUT=gpurun_out
mkdir -p $OUT
run() {  # name, env...
  name=$1; shift
  env "$@" TRAJOPT_B200_TICK_DETAIL=1 TRAJOPT_B200_TICK_LOG=$OUT/${TAG}_ticks_$name.txt timeout 300 python bench.py --batch $BB --steps 1 --warmup 0 --no-cpu-baseline > $OUT/${TAG}_bench_$name.json 2> $OUT/${TAG}_bench_$name.err
  echo "== $name: $(grep 'timed step' $OUT/${TAG}_bench_$name.err)"
  python tools/tick_summary.py $OUT/${TAG}_ticks_$name.txt
}
run base A=1
run jac_pc2_b3 TRAJOPT_B200_JAC_MINB=3
run jac_pc2_b4 TRAJOPT_B200_JAC_MINB=4
run jac_pc1_b2 TRAJOPT_B200_JAC_PC=1
run jac_pc1_b3 TRAJOPT_B200_JAC_PC=1 TRAJOPT_B200_JAC_MINB=3
run jac_pc1_b4_trial4 TRAJOPT_B200_JAC_PC=1 TRAJOPT_B200_JAC_MINB=4 TRAJOPT_B200_TRIAL_MINB=4
BB=16
run tiny A=1
ls $OUT | tail -3
