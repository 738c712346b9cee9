#!/bin/bash
# CTA-per-problem backward pass: parity of every case with the latency path forced on, then tick timing with thresholds
TAG=${1:-r01q}
BB=${2:-16384}
OUT=gpurun_out
mkdir -p $OUT
TRAJOPT_B200_BP_CTA_THRESHOLD=100000000 timeout 600 python tools/gpu_diag.py > $OUT/${TAG}_diag_cta_forced.log 2>&1; echo "diag(cta forced) exit $?"; grep -c "bit-exact" $OUT/${TAG}_diag_cta_forced.log; tail -2 $OUT/${TAG}_diag_cta_forced.log
timeout 600 python tools/gpu_diag.py > $OUT/${TAG}_diag_default.log 2>&1; echo "diag(default) exit $?"; tail -1 $OUT/${TAG}_diag_default.log
run() {
  name=$1; shift
  env "$@" TRAJOPT_B200_TICK_DETAIL=1 TRAJOPT_B200_TICK_LOG=$OUT/${TAG}_ticks_$name.txt timeout 300 python bench.py --batch $BB --steps 1 --warmup 0 --no-cpu-baseline > $OUT/${TAG}_bench_$name.json 2> $OUT/${TAG}_bench_$name.err
  echo "== $name: $(grep 'timed step' $OUT/${TAG}_bench_$name.err)"
  python tools/tick_summary.py $OUT/${TAG}_ticks_$name.txt
}
run off TRAJOPT_B200_BP_CTA_THRESHOLD=0
run default A=1
run t1184 TRAJOPT_B200_BP_CTA_THRESHOLD=1184
run t4096 TRAJOPT_B200_BP_CTA_THRESHOLD=4096
run all TRAJOPT_B200_BP_CTA_THRESHOLD=100000000
