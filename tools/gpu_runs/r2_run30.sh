#!/bin/bash
# round 2, run 30 (two GPUs): does pinning the ranks to disjoint core slices remove the rank skew at the exchanges?
O=gpurun_out/r2_run30; mkdir -p $O
for mode in slice off node; do
  FLOCK_BENCH_BIND=$mode timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 --no-queries > $O/bench_n2_$mode.json 2> $O/bench_n2_$mode.err
  python - $mode <<'PY'
import json, sys
for l in open(f'gpurun_out/r2_run30/bench_n2_{sys.argv[1]}.json'):
    if l.startswith('{'):
        d=json.loads(l); print(sys.argv[1], d['ms_per_step'], d['cpu_affinity'], d['per_rank'], d['queries']['q8']['single_share_ms'])
PY
done
