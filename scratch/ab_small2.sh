#!/bin/bash
line() { python bench.py --workload small 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); c=d['cases']; w=c['wta']; p=c['parity']
print('$1', 'wta fwd', round(w['rk4_forward']['pop_steps_per_sec']/1e9,2), 'fwd+adj', round(w['rk4_forward_adjoint']['pop_steps_per_sec']/1e9,2), 'srk', round(w['srk_forward_adjoint']['pop_steps_per_sec']/1e9,2), '| parity fwd', round(p['rk4_forward']['pop_steps_per_sec']/1e9,2), 'fwd+adj', round(p['rk4_forward_adjoint']['pop_steps_per_sec']/1e9,2))"; }
cp ode-column_b200/lib/libodecol.so /tmp/shipped.so
for l in "$@" "$1"; do cp scratch/libs/$l ode-column_b200/lib/libodecol.so; line $l; done
cp /tmp/shipped.so ode-column_b200/lib/libodecol.so
