#!/bin/bash
# N-GPU confirmation at the full per-GPU batch: torchrun bench (NCCL all_gather of the result records)
TAG=${1:-r01mg4}
NG=${2:-4}
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29527 bench.py --gpus $NG --steps 2 --warmup 1 > $OUT/${TAG}_bench_n$NG.json 2> $OUT/${TAG}_bench_n$NG.err; echo "bench n$NG exit $?"
cut -c1-700 $OUT/${TAG}_bench_n$NG.json; tail -6 $OUT/${TAG}_bench_n$NG.err
