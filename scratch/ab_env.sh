#!/bin/bash
# A/B on ONE box through environment switches: scratch/ab_env.sh "VAR=0" "VAR=1" ... ; the first setting is run again at the end
line() { env $1 python bench.py --steps 3 --warmup 3 --no-secondary --no-cpu-baseline --probe-trials 0 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', round(d['value']/1e9,3), 'e9', round(d['ms_per_step'],1), 'ms; fwd', round(d['phases']['forward_ms'],1), 'bwd', round(d['phases']['loss_backward_plus_adjoint_ms'],1), 'parity', d['parity'], 'clk', d['clocks']['sm_mhz'])"; }
for v in "$@" "$1"; do line "$v"; done
