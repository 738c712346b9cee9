#!/bin/bash
# One GPU-box pass: parity tests, a bench line, the ncu launch list and one full capture of the solve kernel.
# Usage (through gpurun):  bash tools/gpu_round.sh <tag> [bench batch] [ncu batch]
TAG=${1:-r01}
BB=${2:-16384}
NB=${3:-1184}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi > $OUT/${TAG}_nvidia_smi.txt 2>&1
nproc > $OUT/${TAG}_nproc.txt
python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/${TAG}_pytest_gpu.log
tail -3 $OUT/${TAG}_pytest_gpu.log
python bench.py --batch $BB --steps 3 --warmup 3 > $OUT/${TAG}_bench_b${BB}.json 2> $OUT/${TAG}_bench_b${BB}.err; echo "bench exit $?"
cat $OUT/${TAG}_bench_b${BB}.json
python bench.py --batch $NB --steps 1 --warmup 0 --no-cpu-baseline > $OUT/${TAG}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 50 --csv --log-file $OUT/${TAG}_launches.csv \
    python bench.py --batch $NB --steps 1 --warmup 0 --no-cpu-baseline > $OUT/${TAG}_ncu_launches.log 2>&1
echo "ncu launches exit $?"
ncu --set full --clock-control none --import-source on -k regex:solve_kernel -c 1 -f -o $OUT/${TAG}_prof \
    python bench.py --batch $NB --steps 1 --warmup 0 --no-cpu-baseline > $OUT/${TAG}_ncu_full.log 2>&1
echo "ncu full exit $?"
ls -la $OUT
