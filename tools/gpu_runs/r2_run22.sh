#!/bin/bash
# round 2, run 22: the bench line at N = 8 (q8, 1 B events, exchange in the timed region) and the reference arm's N = 8 behaviour
mkdir -p gpurun_out/r2_run22
nvidia-smi topo -m > gpurun_out/r2_run22/topo.txt 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/r2_run22/bench_n8.json 2> gpurun_out/r2_run22/bench_n8.err
tail -c 4000 gpurun_out/r2_run22/bench_n8.json; tail -12 gpurun_out/r2_run22/bench_n8.err
