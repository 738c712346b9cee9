#!/bin/bash
TAG=${1:-r01j}
BB=${2:-16384}
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python tools/gpu_diag.py > $OUT/${TAG}_diag.log 2>&1; echo "diag exit $?"; grep -c "8 / 8" $OUT/${TAG}_diag.log; tail -1 $OUT/${TAG}_diag.log
run() {
  name=$1; shift
  env "$@" TRAJOPT_B200_TICK_DETAIL=1 TRAJOPT_B200_TICK_LOG=$OUT/${TAG}_ticks_$name.txt timeout 300 python bench.py --batch $BB --steps 1 --warmup 0 --no-cpu-baseline > $OUT/${TAG}_bench_$name.json 2> $OUT/${TAG}_bench_$name.err
  echo "== $name: $(grep 'timed step' $OUT/${TAG}_bench_$name.err)"
  python tools/tick_summary.py $OUT/${TAG}_ticks_$name.txt
}
run base A=1
BB=16
run tiny A=1
timeout 120 python bench.py --batch 16 --steps 1 --warmup 0 --no-cpu-baseline > $OUT/${TAG}_plain.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'ls_bp_kernel' -s 30 -c 1 -f -o $OUT/${TAG}_prof_tiny \
    python bench.py --batch 16 --steps 1 --warmup 0 --no-cpu-baseline > $OUT/${TAG}_ncu.log 2>&1
echo "ncu exit $?"
