#!/bin/bash
# End-of-session measurement set on one GPU: tests, smoke, the bench lines and the profiler captures behind profiles/r2_*
set -x
python -m pytest tests -m gpu -q > gpurun_out/final_pytest.log 2>&1; tail -3 gpurun_out/final_pytest.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
python bench.py > gpurun_out/r2_bench_c4.json 2> gpurun_out/r2_bench_c4.err; echo "c4 exit $?"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_bench_c4_reference_arm.json 2> gpurun_out/r2_ref.err; echo "ref exit $?"
python bench.py --workload c5 --c5-horizon 0.05 --steps 1 --warmup 1 > gpurun_out/r2_bench_c5_50ms_1gpu.json 2> gpurun_out/r2_c5.err; echo "c5 exit $?"
python bench.py --workload small > gpurun_out/r2_bench_small.json 2> gpurun_out/r2_small.err; echo "small exit $?"
C4="python bench.py --steps 1 --warmup 1 --time-points 40 --no-secondary --no-cpu-baseline --probe-trials 0 --parity-trials 0"
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r2_launches_c4.csv $C4 > gpurun_out/ncu_c4.log 2>&1
C16="python bench.py --steps 1 --warmup 1 --time-points 16 --no-secondary --no-cpu-baseline --probe-trials 0 --parity-trials 0"
ncu --set full --clock-control none --import-source on -k regex:'k_tc_rk4_fwd_persistent' -s 2 -c 1 -f -o gpurun_out/r2_ncu_fwd $C16 > gpurun_out/ncu_fwd.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_tc_bwd_chain|k_tc_replay' -s 16 -c 4 -f -o gpurun_out/r2_ncu_bwd $C16 > gpurun_out/ncu_bwd.log 2>&1
C5="python bench.py --workload c5 --c5-horizon 0.0004 --steps 1 --warmup 0"
ncu --set full --clock-control none -k regex:'k_tc_contract' -s 20 -c 1 -f -o gpurun_out/r2_ncu_c5 $C5 > gpurun_out/ncu_c5full.log 2>&1
ls -la gpurun_out/r2_ncu_*.ncu-rep
