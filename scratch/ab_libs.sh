#!/bin/bash
# A/B on ONE box with prebuilt libraries: scratch/ab_libs.sh base.so variant1.so ... (each under scratch/libs/); the first is run again at the end (clock drift check)
line() { python bench.py --steps 3 --warmup 3 --no-secondary --no-cpu-baseline --probe-trials 0 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', round(d['value']/1e9,3), 'e9', round(d['ms_per_step'],1), 'ms; fwd', round(d['phases']['forward_ms'],1), 'bwd', round(d['phases']['loss_backward_plus_adjoint_ms'],1), 'parity', d['parity']['ok'], 'clk', d['clocks']['sm_mhz'])"; }
cp ode-column_b200/lib/libodecol.so /tmp/libodecol_shipped.so
for l in "$@" "$1"; do
  cp scratch/libs/$l ode-column_b200/lib/libodecol.so
  line $l
done
cp /tmp/libodecol_shipped.so ode-column_b200/lib/libodecol.so
