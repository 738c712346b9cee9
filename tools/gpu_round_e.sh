#!/bin/bash
# parity diag (tail mode on / off), bench with tick log, ncu of bp + jac variants
TAG=${1:-r01e}
BB=${2:-8192}
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python tools/gpu_diag.py > $OUT/${TAG}_diag_lockstep.log 2>&1; echo "diag exit $?"
grep -c "bit-exact" $OUT/${TAG}_diag_lockstep.log; tail -2 $OUT/${TAG}_diag_lockstep.log
TRAJOPT_B200_TAIL_THRESHOLD=0 timeout 600 python tools/gpu_diag.py quad_altro quad_regdiv cart_altro escape_notebook pend_mintime park_inf_altro > $OUT/${TAG}_diag_notail.log 2>&1; tail -2 $OUT/${TAG}_diag_notail.log
TRAJOPT_B200_TICK_LOG=$OUT/${TAG}_ticks.txt timeout 400 python bench.py --batch $BB --steps 1 --warmup 1 --no-cpu-baseline > $OUT/${TAG}_bench_b${BB}.json 2> $OUT/${TAG}_bench_b${BB}.err; echo "bench exit $?"
cat $OUT/${TAG}_bench_b${BB}.json; tail -4 $OUT/${TAG}_bench_b${BB}.err
for pc in 1 3; do
TRAJOPT_B200_JAC_PC=$pc timeout 300 python bench.py --batch $BB --steps 1 --warmup 1 --no-cpu-baseline > $OUT/${TAG}_bench_pc${pc}.json 2> $OUT/${TAG}_bench_pc${pc}.err; tail -2 $OUT/${TAG}_bench_pc${pc}.err
done
timeout 300 python bench.py --batch $BB --steps 1 --warmup 0 --no-cpu-baseline > $OUT/${TAG}_plain.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'ls_(jac|bp|trial|accept)_kernel' -s 12 -c 6 -f -o $OUT/${TAG}_prof \
    python bench.py --batch $BB --steps 1 --warmup 0 --no-cpu-baseline > $OUT/${TAG}_ncu_full.log 2>&1
echo "ncu full exit $?"
ls -la $OUT | tail -12
