#!/bin/bash
# parity diagnostics of the lockstep engine on every case + a small bench + launch list
TAG=${1:-r01b}
BB=${2:-4096}
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python tools/gpu_diag.py > $OUT/${TAG}_diag_lockstep.log 2>&1; echo "diag exit $?"
tail -5 $OUT/${TAG}_diag_lockstep.log
timeout 300 python bench.py --batch $BB --steps 2 --warmup 1 > $OUT/${TAG}_bench_b${BB}.json 2> $OUT/${TAG}_bench_b${BB}.err; echo "bench exit $?"
cat $OUT/${TAG}_bench_b${BB}.json; tail -15 $OUT/${TAG}_bench_b${BB}.err
TRAJOPT_B200_ENGINE=persistent timeout 300 python bench.py --batch $BB --steps 1 --warmup 1 --no-cpu-baseline > $OUT/${TAG}_bench_persistent_b${BB}.json 2> $OUT/${TAG}_bench_persistent_b${BB}.err; echo "bench persistent exit $?"
cat $OUT/${TAG}_bench_persistent_b${BB}.json; tail -8 $OUT/${TAG}_bench_persistent_b${BB}.err
timeout 200 python bench.py --batch $BB --steps 1 --warmup 0 --no-cpu-baseline > $OUT/${TAG}_plain.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $OUT/${TAG}_launches.csv \
    python bench.py --batch $BB --steps 1 --warmup 0 --no-cpu-baseline > $OUT/${TAG}_ncu_launches.log 2>&1
echo "ncu launches exit $?"
ls -la $OUT
