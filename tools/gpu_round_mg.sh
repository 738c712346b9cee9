#!/bin/bash
# 2-GPU confirmation: parity diag of the working tree, torchrun bench (NCCL all_gather of the result records), reference arm under torchrun
TAG=${1:-r01mg}
BB=${2:-8192}
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python tools/gpu_diag.py > $OUT/${TAG}_diag.log 2>&1; echo "diag exit $?"; tail -3 $OUT/${TAG}_diag.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --batch $BB --steps 1 --warmup 1 > $OUT/${TAG}_bench_n2.json 2> $OUT/${TAG}_bench_n2.err; echo "bench n2 exit $?"
cut -c1-600 $OUT/${TAG}_bench_n2.json; tail -5 $OUT/${TAG}_bench_n2.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --impl reference --gpus 2 --steps 1 --warmup 1 > $OUT/${TAG}_ref_n2.json 2> $OUT/${TAG}_ref_n2.err; echo "ref n2 exit $?"
cut -c1-300 $OUT/${TAG}_ref_n2.json
