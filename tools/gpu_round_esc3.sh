#!/bin/bash
TAG=${1:-r01e8}
OUT=gpurun_out
mkdir -p $OUT
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'ls_trial_kernel' -s 300 -c 1 -f -o $OUT/${TAG}_prof_escape_trial \
    python tools/run_configs.py ${TAG} 64 escape_altro > $OUT/${TAG}_ncu_escape.log 2>&1
echo "ncu exit $?"
TRAJOPT_B200_TICK_DETAIL=1 TRAJOPT_B200_TICK_LOG=$OUT/${TAG}_ticks_acrobot_sqrt_mt.txt timeout 600 python tools/run_configs.py ${TAG}a 1 acrobot_sqrt_mintime > $OUT/${TAG}_acrobot.log 2>&1
python tools/tick_summary.py $OUT/${TAG}_ticks_acrobot_sqrt_mt.txt
