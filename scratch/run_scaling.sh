#!/bin/bash
# bench lines at N GPUs of one box: C4 (default line incl. probe + C5 secondary) and C5 at the 0.05 s horizon
N=$1
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N "$@"; }
run --steps 3 --warmup 3 > gpurun_out/r2_bench_c4_${N}gpu.json 2> gpurun_out/r2_bench_c4_${N}gpu.err; echo "c4 exit $?"
run --workload c5 --c5-horizon 0.05 --steps 1 --warmup 1 > gpurun_out/r2_bench_c5_50ms_${N}gpu.json 2> gpurun_out/r2_bench_c5_${N}gpu.err; echo "c5 exit $?"
python - <<PY
import json
for f in ("gpurun_out/r2_bench_c4_${N}gpu.json", "gpurun_out/r2_bench_c5_50ms_${N}gpu.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d["n_gpus"], d["value"], d["ms_per_step"], d.get("probe"), (d.get("secondary") or {}).get("value"))
    except Exception as e:
        print(f, "unreadable", e)
PY
