#!/bin/bash
# round 2, run 31 (two GPUs): per-rank kernel profiles of a q8 step -- which rank reaches the exchanges late, and in which kernels
O=gpurun_out/r2_run31; mkdir -p $O
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 --no-queries > $O/bench_n2.json 2> $O/bench_n2.err
python - <<'PY'
import json
for l in open('gpurun_out/r2_run31/bench_n2.json'):
    if l.startswith('{'):
        d=json.loads(l); print(d['ms_per_step'])
        ks=sorted(set(k for r in d['per_rank'] for k in r['kernels_us']))
        for k in ks: print(f"{k:34s}", [r['kernels_us'].get(k) for r in d['per_rank']])
PY
