#!/bin/bash
TAG=${1:-r01i}
OUT=gpurun_out
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -3 $OUT/${TAG}_pytest_gpu.log
timeout 120 python __graft_entry__.py smoke > $OUT/${TAG}_smoke.log 2>&1; echo "smoke exit $?"; tail -2 $OUT/${TAG}_smoke.log
timeout 900 python bench.py > $OUT/${TAG}_bench_default.json 2> $OUT/${TAG}_bench_default.err; echo "bench exit $?"
cat $OUT/${TAG}_bench_default.json; tail -12 $OUT/${TAG}_bench_default.err
