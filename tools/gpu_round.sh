#!/bin/bash
# One parameterised GPU round (replaces the per-experiment gpu_round_*.sh scripts).
#   tools/gpu_round.sh TAG step [step...]      steps: test testk=EXPR smoke bench bench16k ref configs cfg=CASE ncu_launches ncu_full ticks16k env=VAR=VALUE tag=TAG
# Everything lands in gpurun_out/TAG_*.  Bench numbers are never taken under ncu.
TAG=$1; shift
OUT=gpurun_out
mkdir -p $OUT
for step in "$@"; do
  case $step in
    test)   timeout 1800 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -3 $OUT/${TAG}_pytest_gpu.log ;;
    testk=*) timeout 1800 python -m pytest tests -m gpu -x -q -k "${step#testk=}" > $OUT/${TAG}_pytest_gpu_k.log 2>&1; echo "pytest -k exit $?"; tail -3 $OUT/${TAG}_pytest_gpu_k.log ;;
    smoke)  timeout 300 python __graft_entry__.py smoke > $OUT/${TAG}_smoke.log 2>&1; echo "smoke exit $?"; tail -2 $OUT/${TAG}_smoke.log ;;
    bench)  timeout 900 python bench.py > $OUT/${TAG}_bench_default.json 2> $OUT/${TAG}_bench_default.err; echo "bench exit $?"
            cut -c1-600 $OUT/${TAG}_bench_default.json; tail -4 $OUT/${TAG}_bench_default.err ;;
    bench16k) TRAJOPT_B200_TICK_LOG=$OUT/${TAG}_ticks16k.txt TRAJOPT_B200_TICK_DETAIL=1 timeout 600 python bench.py --batch 16384 --steps 2 --warmup 1 --no-cpu-baseline \
              > $OUT/${TAG}_bench16k.json 2> $OUT/${TAG}_bench16k.err; echo "bench16k exit $?"; cut -c1-300 $OUT/${TAG}_bench16k.json; tail -3 $OUT/${TAG}_bench16k.err ;;
    ref)    timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/${TAG}_bench_reference.json 2> $OUT/${TAG}_bench_reference.err; echo "reference exit $?"
            cut -c1-300 $OUT/${TAG}_bench_reference.json ;;
    configs) timeout 1500 python tools/run_configs.py ${TAG} 1 > $OUT/${TAG}_configs.log 2>&1; echo "configs exit $?"; tail -12 $OUT/${TAG}_configs.log | cut -c1-400 ;;
    configs8) timeout 900 python tools/run_configs.py ${TAG}s8 8 > $OUT/${TAG}_configs8.log 2>&1; echo "configs/8 exit $?"; tail -12 $OUT/${TAG}_configs8.log | cut -c1-400 ;;
    ncu_launches)
            timeout 300 python bench.py --batch 8192 --steps 1 --warmup 0 --no-cpu-baseline > $OUT/${TAG}_plain.log 2>&1 &&
            timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 1 -c 640 --csv --log-file $OUT/${TAG}_launches.csv \
                python bench.py --batch 8192 --steps 1 --warmup 0 --no-cpu-baseline > $OUT/${TAG}_ncu_launches.log 2>&1
            echo "ncu launches exit $?" ;;
    ncu_full)
            timeout 300 python bench.py --batch 8192 --steps 1 --warmup 0 --no-cpu-baseline > $OUT/${TAG}_plain2.log 2>&1 &&
            timeout 900 ncu --set full --clock-control none --import-source on -k regex:'ls_(jac|bp|trial|accept_tail)_kernel' -s 12 -c 6 -f -o $OUT/${TAG}_prof \
                python bench.py --batch 8192 --steps 1 --warmup 0 --no-cpu-baseline > $OUT/${TAG}_ncu_full.log 2>&1
            echo "ncu full exit $?" ;;
    ncu_resident)
            timeout 300 python bench.py --batch 64 --steps 1 --warmup 0 --no-cpu-baseline > $OUT/${TAG}_plain3.log 2>&1 &&
            timeout 900 ncu --set full --clock-control none --import-source on -k regex:'ls_resident_kernel' -c 1 -f -o $OUT/${TAG}_prof_resident \
                python bench.py --batch 64 --steps 1 --warmup 0 --no-cpu-baseline > $OUT/${TAG}_ncu_resident.log 2>&1
            echo "ncu resident exit $?" ;;
    resprof) for c in quad_altro escape_notebook cart_altro; do timeout 300 python tools/resident_profile.py $c 8; done > $OUT/${TAG}_resident_profile.log 2>&1
            echo "resprof exit $?"; cat $OUT/${TAG}_resident_profile.log ;;
    ncu_resprof=*) RCASE="${step#ncu_resprof=}"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:ls_resident_kernel -c 1 -f -o $OUT/${TAG}_prof_resident python tools/resident_profile.py $RCASE 8 > $OUT/${TAG}_ncu_resident.log 2>&1; echo "ncu resident ($RCASE) exit $?" ;;
    ncu_resprof)
            timeout 900 ncu --set full --clock-control none --import-source on -k regex:'ls_resident_kernel' -c 1 -f -o $OUT/${TAG}_prof_resident \
                python tools/resident_profile.py quad_altro 8 > $OUT/${TAG}_ncu_resident.log 2>&1
            echo "ncu resident exit $?" ;;
    cfg=*)  CNAME="${step#cfg=}"; TRAJOPT_B200_TICK_LOG=$OUT/${TAG}_ticks_${CNAME}.txt TRAJOPT_B200_TICK_DETAIL=1 timeout 900 python tools/run_configs.py ${TAG}_${CNAME} 1 $CNAME > $OUT/${TAG}_cfg_${CNAME}.log 2>&1
            echo "config $CNAME exit $?"; tail -2 $OUT/${TAG}_cfg_${CNAME}.log | cut -c1-500; tail -4 $OUT/${TAG}_ticks_${CNAME}.txt ;;
    sqrttime) timeout 600 python tools/sqrt_pass_timing.py 8 > $OUT/${TAG}_sqrt_pass_timing.log 2>&1; echo "sqrttime exit $?"; cat $OUT/${TAG}_sqrt_pass_timing.log ;;
    bpprof) timeout 300 python tools/bp_profile.py 8192 > $OUT/${TAG}_bp_profile.log 2>&1; echo "bpprof exit $?"; cat $OUT/${TAG}_bp_profile.log ;;
    env=*)  export "${step#env=}" ;;
    tag=*)  TAG="${step#tag=}" ;;
    *) echo "unknown step $step" ;;
  esac
done
ls -la $OUT | tail -5
