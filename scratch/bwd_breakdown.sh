#!/bin/bash
# where the reverse sweep's time goes: diagnostics build (-DODECOL_DIAG), one box, phases of the C4 pass under the skip masks
ODECOL_NVCC_EXTRA="-DODECOL_DIAG" python ode-column_b200/build.py --force > /tmp/build.log 2>&1 || tail -5 /tmp/build.log
line() { python bench.py --steps 2 --warmup 2 --no-secondary --no-cpu-baseline --probe-trials 0 --parity-trials 0 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', 'fwd', round(d['phases']['forward_ms'],1), 'bwd', round(d['phases']['loss_backward_plus_adjoint_ms'],1), 'launches', d['gpu_launches'], 'clk', d['clocks']['sm_mhz'])"; }
line default
ODECOL_OVERLAP=0 line overlap0
ODECOL_DBG_BWD_SKIP=1 line skip_dw
ODECOL_DBG_BWD_SKIP=2 line skip_chain
ODECOL_DBG_BWD_SKIP=3 line replay_only
ODECOL_DBG_BWD_SKIP=3 ODECOL_OVERLAP=0 line replay_only_serial
ODECOL_FUSE_DW=1 line fuse_dw
ODECOL_PERSISTENT=0 line per_stage_launches
