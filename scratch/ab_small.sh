#!/bin/bash
# A/B of the on-chip family on ONE box: shipped build vs a rebuild with extra nvcc flags ($1)
line() { python bench.py --workload small --steps 3 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); c=d['cases']
print('$1', 'wta fwd %.3ge10 fwd+adj %.3ge9 srk %.3ge9 | parity fwd %.3ge9 fwd+adj %.3ge9 | clk %s' % (c['wta']['rk4_forward']['pop_steps_per_sec']/1e10, c['wta']['rk4_forward_adjoint']['pop_steps_per_sec']/1e9, c['wta']['srk_forward_adjoint']['pop_steps_per_sec']/1e9, c['parity']['rk4_forward']['pop_steps_per_sec']/1e9, c['parity']['rk4_forward_adjoint']['pop_steps_per_sec']/1e9, d['clocks']['sm_mhz']))"; }
line A
cp ode-column_b200/lib/libodecol.so /tmp/libodecol_A.so
ODECOL_NVCC_EXTRA="$1" python ode-column_b200/build.py --force > /tmp/build.log 2>&1 || tail -5 /tmp/build.log
line "B[$1]"
python -m pytest -m gpu -q tests/test_gpu_parity.py -k "golden or trials_are_independent or packing" 2>&1 | tail -3
cp /tmp/libodecol_A.so ode-column_b200/lib/libodecol.so
