#!/bin/bash
# A/B on ONE box: the shipped build, a rebuild with extra nvcc flags ($1), the shipped build again (clock drift check)
line() { python bench.py --steps 3 --warmup 3 --no-secondary --no-cpu-baseline --probe-trials 0 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', round(d['value']/1e9,3), 'e9', round(d['ms_per_step'],1), 'ms; fwd', round(d['phases']['forward_ms'],1), 'bwd', round(d['phases']['loss_backward_plus_adjoint_ms'],1), 'parity', d['parity']['ok'], 'clk', d['clocks']['sm_mhz'])"; }
line A
cp ode-column_b200/lib/libodecol.so /tmp/libodecol_A.so
ODECOL_NVCC_EXTRA="$1" python ode-column_b200/build.py --force > /tmp/build.log 2>&1 || tail -5 /tmp/build.log
line "B[$1]"
cp /tmp/libodecol_A.so ode-column_b200/lib/libodecol.so
line A
