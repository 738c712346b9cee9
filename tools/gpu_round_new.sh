#!/bin/bash
TAG=${1:-r01n2}
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python tools/gpu_diag.py car_3obs_altro quad_obs_al > $OUT/${TAG}_diag.log 2>&1; tail -4 $OUT/${TAG}_diag.log
TRAJOPT_B200_TAIL_THRESHOLD=0 TRAJOPT_B200_BP_CTA_THRESHOLD=0 timeout 900 python tools/gpu_diag.py car_3obs_altro quad_obs_al > $OUT/${TAG}_diag_bulk.log 2>&1; tail -4 $OUT/${TAG}_diag_bulk.log
TRAJOPT_B200_ENGINE=persistent timeout 900 python tools/gpu_diag.py car_3obs_altro quad_obs_al > $OUT/${TAG}_diag_persistent.log 2>&1; tail -4 $OUT/${TAG}_diag_persistent.log
