#!/bin/bash
# round 2, run 38 (two GPUs): does the NVML clock sampler perturb the host-driven q8 step?  2 ms vs 10 ms sampling; then the two-GPU tests on the final code
O=gpurun_out/r2_run38; mkdir -p $O
for ms in 2 10 2 10; do
  FLOCK_BENCH_SAMPLE_MS=$ms timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 40 --warmup 3 --no-queries > $O/bench_n2_$ms.json 2> $O/bench_n2_$ms.err
  python - $ms <<'PY'
import json, sys
for l in open(f'gpurun_out/r2_run38/bench_n2_{sys.argv[1]}.json'):
    if l.startswith('{'):
        d=json.loads(l); print('sample every', sys.argv[1], 'ms:', round(d['ms_per_step'],4), 'alone', round(d['queries']['q8']['single_share_ms'],4), 'e2e ms', round(d['e2e']['ms_per_step'],3), d['clocks'])
PY
done
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q > $O/multi.log 2>&1; tail -3 $O/multi.log
