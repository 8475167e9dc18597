#!/bin/bash
# round 2, run 32 (eight GPUs): the bench line at N = 8 after the scatter / probe fixes, per-rank kernel profiles; the same with the NCCL fallback for comparison
O=gpurun_out/r2_run32; mkdir -p $O
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 20 --warmup 3 > $O/bench_n8.json 2> $O/bench_n8.err
python - <<'PY'
import json
for l in open('gpurun_out/r2_run32/bench_n8.json'):
    if l.startswith('{'):
        d=json.loads(l); print({k:d[k] for k in ('value','ms_per_step','phases_ms','launches_per_step_rank0','parity_check','cpu_affinity') if k in d}); print(d['queries'])
        ks=sorted(set(k for r in d['per_rank'] for k in r['kernels_us']))
        for k in ks: print(f"{k:34s}", [r['kernels_us'].get(k) for r in d['per_rank']])
PY
tail -2 $O/bench_n8.err
FLOCKGPU_EXCHANGE=nccl timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 8 --steps 10 --warmup 3 --no-queries > $O/bench_n8_nccl.json 2> $O/bench_n8_nccl.err
python - <<'PY'
import json
for l in open('gpurun_out/r2_run32/bench_n8_nccl.json'):
    if l.startswith('{'):
        d=json.loads(l); print('NCCL fallback:', {k:d[k] for k in ('ms_per_step','phases_ms','launches_per_step_rank0','parity_check') if k in d}); print({k:round(v['ms'],4) for k,v in d['kernels'].items()})
PY
tail -2 $O/bench_n8_nccl.err
