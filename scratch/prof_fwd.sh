#!/bin/bash
# ncu --set full capture of the persistent forward kernel (C4, 16 grid points) with source
CMD="python bench.py --steps 1 --warmup 1 --time-points 16 --no-secondary --no-cpu-baseline --probe-trials 0 --parity-trials 0"
$CMD > gpurun_out/prof_fwd_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/prof_fwd_plain.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:'k_tc_rk4_fwd_persistent' -s 2 -c 1 -f -o gpurun_out/prof_fwd_$1 $CMD > gpurun_out/prof_fwd_ncu.log 2>&1
tail -2 gpurun_out/prof_fwd_ncu.log
