#!/bin/bash
# round 2, run 33 (four GPUs): the bench line at N = 4
O=gpurun_out/r2_run33; mkdir -p $O
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 4 --steps 20 --warmup 3 > $O/bench_n4.json 2> $O/bench_n4.err
python - <<'PY'
import json
for l in open('gpurun_out/r2_run33/bench_n4.json'):
    if l.startswith('{'):
        d=json.loads(l); print({k:d[k] for k in ('value','ms_per_step','phases_ms','launches_per_step_rank0','cpu_affinity') if k in d}); print(d['parity_check']['q8']); print(d['queries'])
        ks=sorted(set(k for r in d['per_rank'] for k in r['kernels_us']))
        for k in ks: print(f"{k:34s}", [r['kernels_us'].get(k) for r in d['per_rank']])
PY
tail -2 $O/bench_n4.err
