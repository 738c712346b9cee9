#!/bin/bash
# restart-chain hand-over (bulk backward pass -> CTA kernel): parity with the bulk kernels forced on, then timing with / without
TAG=${1:-r01d2}
OUT=gpurun_out
mkdir -p $OUT
TRAJOPT_B200_TAIL_THRESHOLD=0 TRAJOPT_B200_BP_CTA_THRESHOLD=0 timeout 900 python tools/gpu_diag.py > $OUT/${TAG}_diag_bulk_defer.log 2>&1; tail -1 $OUT/${TAG}_diag_bulk_defer.log
timeout 600 python tools/gpu_diag.py quad_altro quad_regdiv quad_obs_al > $OUT/${TAG}_diag_default.log 2>&1; tail -1 $OUT/${TAG}_diag_default.log
timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "kernels_agree or full_size or status_bits" 2>&1 | tail -2
run() {
  name=$1; bb=$2; shift; shift
  env "$@" TRAJOPT_B200_TICK_DETAIL=1 TRAJOPT_B200_TICK_LOG=$OUT/${TAG}_ticks_$name.txt timeout 300 python bench.py --batch $bb --steps 1 --warmup 0 --no-cpu-baseline > $OUT/${TAG}_bench_$name.json 2> $OUT/${TAG}_bench_$name.err
  echo "== $name: $(grep 'timed step' $OUT/${TAG}_bench_$name.err)"
  python tools/tick_summary.py $OUT/${TAG}_ticks_$name.txt
}
run defer16k 16384 A=1
run nodefer16k 16384 TRAJOPT_B200_BP_DEFER_RESTARTS=0
run defer64k 65536 A=1
run nodefer64k 65536 TRAJOPT_B200_BP_DEFER_RESTARTS=0
