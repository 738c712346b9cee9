#!/bin/bash
# square-root backward pass after the knot-parallel expansion + unrolled recursion: parity of every sqrt case, full-size timing
TAG=${1:-r01sq}
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python tools/gpu_diag.py pend_sqrt_altro dp_sqrt_ilqr acrobot_sqrt_al acrobot_sqrt_mintime dp_sqrt_mintime > $OUT/${TAG}_diag.log 2>&1; grep -E "^==|bit-exact|TOTAL" $OUT/${TAG}_diag.log | cut -c1-150
timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "sqrt or status_bits" 2>&1 | tail -2
TRAJOPT_B200_TICK_DETAIL=1 TRAJOPT_B200_TICK_LOG=$OUT/${TAG}_ticks_acrobot_sqrt_mt.txt timeout 600 python tools/run_configs.py ${TAG} 1 acrobot_sqrt_mintime dp_sqrt_mintime > $OUT/${TAG}_configs.log 2>&1
cut -c1-330 $OUT/${TAG}_configs.log
python tools/tick_summary.py $OUT/${TAG}_ticks_acrobot_sqrt_mt.txt
