#!/bin/bash
# micro-benchmarks + full ncu capture of one tick of the lockstep kernels
TAG=${1:-r01c}
BB=${2:-8192}
OUT=gpurun_out
mkdir -p $OUT
timeout 120 ./tools/fp64_micro > $OUT/${TAG}_fp64_micro.log 2>&1; echo "micro exit $?"
cat $OUT/${TAG}_fp64_micro.log
timeout 200 python bench.py --batch $BB --steps 1 --warmup 0 --no-cpu-baseline > $OUT/${TAG}_plain.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'ls_(jac|bp|trial|accept)_kernel' -s 12 -c 6 -f -o $OUT/${TAG}_prof \
    python bench.py --batch $BB --steps 1 --warmup 0 --no-cpu-baseline > $OUT/${TAG}_ncu_full.log 2>&1
echo "ncu full exit $?"
tail -5 $OUT/${TAG}_ncu_full.log
ls -la $OUT
