#!/bin/bash
# ncu --set full capture of the C4 tensor-family kernels (forward persistent, replay, reverse-stage chain, dW) with source
CMD="python bench.py --steps 1 --warmup 1 --time-points 16 --no-secondary --no-cpu-baseline --probe-trials 0 --parity-trials 0"
$CMD > gpurun_out/prof_s4_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/prof_s4_plain.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:'k_tc_rk4_fwd_persistent|k_tc_bwd_chain|k_tc_replay|k_tc_dw' -s 8 -c 8 -f -o gpurun_out/prof_s4_c4 $CMD > gpurun_out/prof_s4_ncu.log 2>&1
tail -3 gpurun_out/prof_s4_ncu.log
ls -la gpurun_out/prof_s4_c4.ncu-rep
