#!/bin/bash
# round 2, run 21: routing of nullable keys agrees across ranks, Final DISTINCT through the one-pass kernels, per-rank exchange waits in the bench line
mkdir -p gpurun_out/r2_run21
timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -q -k "nulls or unknown_types or unique_probe or partition or aggregate" > gpurun_out/r2_run21/focus.log 2>&1
tail -15 gpurun_out/r2_run21/focus.log
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_nexmark.py -m gpu -q > gpurun_out/r2_run21/multi.log 2>&1
tail -40 gpurun_out/r2_run21/multi.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r2_run21/bench_n2.json 2> gpurun_out/r2_run21/bench_n2.err
tail -c 3000 gpurun_out/r2_run21/bench_n2.json; tail -5 gpurun_out/r2_run21/bench_n2.err
