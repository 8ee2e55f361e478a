#!/bin/bash
# forward time per (tile, stage pass) with 1, 2, 3, 4 trial tiles per CTA: separates "dependency chain per tile" from "epilogue throughput"
for tr in 4144 8288 12432 16576; do
python bench.py --steps 2 --warmup 1 --time-points 200 --trials-per-gpu $tr --no-secondary --no-cpu-baseline --probe-trials 0 --parity-trials 0 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); f=d['phases']['forward_ms']; b=d['phases']['loss_backward_plus_adjoint_ms']; tiles=$tr/112*4/148
print('trials', $tr, 'tiles/CTA', tiles, 'fwd ms', round(f,2), 'us per tile-stage', round(f*1e3/(199*4*tiles),2), '| bwd ms', round(b,2), 'us per tile-step', round(b*1e3/(199*tiles),2), 'clk', d['clocks']['sm_mhz'])"
done
