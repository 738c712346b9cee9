#!/bin/bash
# end-of-round evidence: GPU tests, smoke, the default bench line, ncu launch list + full capture of one tick
TAG=${1:-r01z}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,power.limit --format=csv > $OUT/${TAG}_gpu.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -2 $OUT/${TAG}_pytest_gpu.log
timeout 120 python __graft_entry__.py smoke > $OUT/${TAG}_smoke.log 2>&1; echo "smoke exit $?"; tail -2 $OUT/${TAG}_smoke.log
timeout 900 python bench.py > $OUT/${TAG}_bench_default.json 2> $OUT/${TAG}_bench_default.err; echo "bench exit $?"
cut -c1-400 $OUT/${TAG}_bench_default.json; tail -4 $OUT/${TAG}_bench_default.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/${TAG}_bench_reference.json 2> $OUT/${TAG}_bench_reference.err; echo "reference exit $?"
cut -c1-300 $OUT/${TAG}_bench_reference.json
timeout 300 python bench.py --batch 8192 --steps 1 --warmup 0 --no-cpu-baseline > $OUT/${TAG}_plain.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 1 -c 640 --csv --log-file $OUT/${TAG}_launches.csv \
    python bench.py --batch 8192 --steps 1 --warmup 0 --no-cpu-baseline > $OUT/${TAG}_ncu_launches.log 2>&1
echo "ncu launches exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'ls_(jac|bp|trial|accept_tail)_kernel' -s 12 -c 6 -f -o $OUT/${TAG}_prof \
    python bench.py --batch 8192 --steps 1 --warmup 0 --no-cpu-baseline > $OUT/${TAG}_ncu_full.log 2>&1
echo "ncu full exit $?"
# the latency path of the backward pass (expansion kernel + CTA per problem): one tick with 1,024 live problems
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'ls_(expand|bp_cta)_kernel' -s 10 -c 2 -f -o $OUT/${TAG}_prof_cta \
    python bench.py --batch 1024 --steps 1 --warmup 0 --no-cpu-baseline > $OUT/${TAG}_ncu_full_cta.log 2>&1
echo "ncu full (cta) exit $?"
ls -la $OUT | tail -8
