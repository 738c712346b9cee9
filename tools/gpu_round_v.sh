#!/bin/bash
# variant sweep (env-selected kernels) with per-phase tick timing: name=ENV... pairs after the tag
TAG=${1:-r01v}
BB=${2:-16384}
OUT=gpurun_out
mkdir -p $OUT
run() {
  name=$1; shift
  env "$@" TRAJOPT_B200_TICK_DETAIL=1 TRAJOPT_B200_TICK_LOG=$OUT/${TAG}_ticks_$name.txt timeout 300 python bench.py --batch $BB --steps 1 --warmup 0 --no-cpu-baseline > $OUT/${TAG}_bench_$name.json 2> $OUT/${TAG}_bench_$name.err
  echo "== $name: $(grep 'timed step' $OUT/${TAG}_bench_$name.err)"
  python tools/tick_summary.py $OUT/${TAG}_ticks_$name.txt | head -2
}
run pc1b3 A=1
run pc2b2 TRAJOPT_B200_JAC_PC=2 TRAJOPT_B200_JAC_MINB=2
run pc2b3 TRAJOPT_B200_JAC_PC=2 TRAJOPT_B200_JAC_MINB=3
timeout 300 python tools/gpu_diag.py quad_altro quad_ilqr quad_regdiv > $OUT/${TAG}_diag_pc1.log 2>&1; tail -1 $OUT/${TAG}_diag_pc1.log
TRAJOPT_B200_JAC_PC=2 TRAJOPT_B200_JAC_MINB=3 timeout 300 python tools/gpu_diag.py quad_altro quad_ilqr quad_regdiv > $OUT/${TAG}_diag_pc2.log 2>&1; tail -1 $OUT/${TAG}_diag_pc2.log
