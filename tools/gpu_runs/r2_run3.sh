#!/bin/bash
# round 2, run 3 (2 GPUs): peer-window exchange + NCCL fallback parity, then q8 sharded timings
mkdir -p gpurun_out/r2_run3
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/r2_run3/pytest_multi.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_run3/pytest_multi.log
tail -30 gpurun_out/r2_run3/pytest_multi.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/nexmark_bench.py --queries q8,q3,q5 --q8-scale 0.25 --reps 10 --no-cpu --no-e2e > gpurun_out/r2_run3/nexmark_2gpu.jsonl 2> gpurun_out/r2_run3/nexmark_2gpu.err
tail -5 gpurun_out/r2_run3/nexmark_2gpu.err
