#!/bin/bash
# end-of-round verification of HEAD without profiler passes: GPU tests, smoke, the default bench line, the reference arm
TAG=${1:-r01fin}
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -2 $OUT/${TAG}_pytest_gpu.log
timeout 120 python __graft_entry__.py smoke > $OUT/${TAG}_smoke.log 2>&1; echo "smoke exit $?"; tail -2 $OUT/${TAG}_smoke.log
timeout 600 python bench.py > $OUT/${TAG}_bench_default.json 2> $OUT/${TAG}_bench_default.err; echo "bench exit $?"
cut -c1-300 $OUT/${TAG}_bench_default.json; tail -3 $OUT/${TAG}_bench_default.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/${TAG}_bench_reference.json 2> $OUT/${TAG}_bench_reference.err; echo "reference exit $?"
cut -c1-200 $OUT/${TAG}_bench_reference.json
