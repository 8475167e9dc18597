#!/bin/bash
# round 2, run 23: peer-store microbenchmark (store width x segment length), host view of a q8 step, N = 2 bench with NUMA binding
mkdir -p gpurun_out/r2_run23
timeout 300 ./build/microbench_p2p > gpurun_out/r2_run23/microbench_p2p.txt 2>&1
cat gpurun_out/r2_run23/microbench_p2p.txt
FLOCKGPU_HOST_TRACE=1 timeout 300 python tools/diag.py q8 > gpurun_out/r2_run23/diag_q8.txt 2>&1
tail -30 gpurun_out/r2_run23/diag_q8.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r2_run23/bench_n2.json 2> gpurun_out/r2_run23/bench_n2.err
python - <<'PY'
import json
for l in open('gpurun_out/r2_run23/bench_n2.json'):
    if l.startswith('{'):
        d=json.loads(l); print({k:d[k] for k in ('value','ms_per_step','phases_ms','per_rank','cpu_affinity','launches_per_step_rank0','parity_check') if k in d}); print(d['kernels']); print(d['queries']['q8'])
PY
tail -3 gpurun_out/r2_run23/bench_n2.err
