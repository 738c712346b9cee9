#!/bin/bash
# parity diag (both engines on the divergent case), bench at BB with both timing legs, ncu of one tick
TAG=${1:-r01d}
BB=${2:-8192}
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python tools/gpu_diag.py > $OUT/${TAG}_diag_lockstep.log 2>&1; echo "diag exit $?"
grep -c "bit-exact" $OUT/${TAG}_diag_lockstep.log; tail -3 $OUT/${TAG}_diag_lockstep.log
TRAJOPT_B200_ENGINE=persistent timeout 300 python tools/gpu_diag.py quad_regdiv quad_altro di_altro > $OUT/${TAG}_diag_persistent.log 2>&1; tail -3 $OUT/${TAG}_diag_persistent.log
timeout 400 python bench.py --batch $BB --steps 2 --warmup 1 > $OUT/${TAG}_bench_b${BB}.json 2> $OUT/${TAG}_bench_b${BB}.err; echo "bench exit $?"
cat $OUT/${TAG}_bench_b${BB}.json; tail -8 $OUT/${TAG}_bench_b${BB}.err
timeout 300 python bench.py --batch $BB --steps 1 --warmup 0 --no-cpu-baseline > $OUT/${TAG}_plain.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'ls_(jac|bp|trial|accept)_kernel' -s 12 -c 6 -f -o $OUT/${TAG}_prof \
    python bench.py --batch $BB --steps 1 --warmup 0 --no-cpu-baseline > $OUT/${TAG}_ncu_full.log 2>&1
echo "ncu full exit $?"
tail -3 $OUT/${TAG}_ncu_full.log
ls -la $OUT | tail -12
