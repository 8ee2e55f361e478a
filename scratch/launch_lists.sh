#!/bin/bash
# ncu launch lists (gpu__time_duration.sum, --clock-control none) of the two headline workloads, after a plain run of each
C4="python bench.py --steps 1 --warmup 1 --time-points 40 --no-secondary --no-cpu-baseline --probe-trials 0 --parity-trials 0"
C5="python bench.py --workload c5 --c5-horizon 0.0004 --steps 1 --warmup 0"
$C4 > gpurun_out/plain_c4.log 2>&1 || { echo C4 plain failed; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r2_launches_c4.csv $C4 > gpurun_out/ncu_c4.log 2>&1
$C5 > gpurun_out/plain_c5.log 2>&1 || { echo C5 plain failed; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_c5.csv $C5 > gpurun_out/ncu_c5.log 2>&1
echo done
