#!/bin/bash
# round 2, run 15 (2 GPUs): TMA ring A/B on q5, multi-GPU tests, bench --gpus 2
mkdir -p gpurun_out/r2_run15
for t in 1 0; do
  echo "== FLOCKGPU_HIST_TMA=$t"
  FLOCKGPU_HIST_TMA=$t timeout 200 python tools/diag.py q5 > gpurun_out/r2_run15/diag_q5_tma$t.txt 2>&1
  head -1 gpurun_out/r2_run15/diag_q5_tma$t.txt | cut -c1-170; grep agg_ gpurun_out/r2_run15/diag_q5_tma$t.txt
done
timeout 600 python -m pytest tests/test_gpu_ops.py tests/test_gpu_baseline_sizes.py -m gpu -x -q -k "dense or q5 or q8 or aggregate" > gpurun_out/r2_run15/pytest_1gpu.log 2>&1
tail -3 gpurun_out/r2_run15/pytest_1gpu.log
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/r2_run15/pytest_multi.log 2>&1
tail -5 gpurun_out/r2_run15/pytest_multi.log
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 50 --warmup 5 --no-queries ) > gpurun_out/r2_run15/bench_n2.json 2> gpurun_out/r2_run15/bench_n2.err
tail -4 gpurun_out/r2_run15/bench_n2.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2_run15/bench_n2.json') if l.startswith('{')][0])
print('q8 x2: ms/step', d['ms_per_step'], 'single share', d['notes']['single_share_ms_per_step'], 'launches/step', d['launches_per_step_rank0'], d['parity_check']['q8'])
print(d['phases_ms'])
print({k:(v['launches'], round(v['ms'],4)) for k,v in d['kernels'].items()})
PY
