#!/bin/bash
# round 2, run 29 (two GPUs): scatter with coalesced offsets stores; partition / exchange tests in both modes; bench N = 2 and its reference arm
O=gpurun_out/r2_run29; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_ops.py -m gpu -q -k "multi or two_gpu or partition or sort or nulls" > $O/focus.log 2>&1; tail -4 $O/focus.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > $O/bench_n2.json 2> $O/bench_n2.err
python - <<'PY'
import json
for l in open('gpurun_out/r2_run29/bench_n2.json'):
    if l.startswith('{'):
        d=json.loads(l); print({k:d[k] for k in ('value','ms_per_step','phases_ms','per_rank','launches_per_step_rank0','parity_check') if k in d}); print({k:round(v['ms'],4) for k,v in d['kernels'].items()}); print(d['queries'])
PY
tail -3 $O/bench_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 5 --warmup 1 > $O/bench_ref_n2.json 2> $O/bench_ref_n2.err; tail -c 600 $O/bench_ref_n2.json
