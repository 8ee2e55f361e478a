#!/bin/bash
# C5 (adaptive Euler-Maruyama sweep, N = 8192) A/B through environment switches
line() { env $1 python bench.py --workload c5 --c5-horizon 0.004 --steps 1 --warmup 1 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', round(d['value']/1e9,3), 'e9', round(d['ms_per_step'],1), 'ms; rounds', d.get('rounds_per_solve'), 'attempted', d.get('attempted_steps_per_member'), 'accepted', d.get('accepted_steps_per_member'), 'finite', d.get('all_members_finite'), 'roofline', round(d['roofline']['frac'],3), 'clk', d['clocks']['sm_mhz'])"; }
for v in "$@" "$1"; do line "$v"; done
