#!/bin/bash
# round 2, run 10 (2 GPUs): q5 after the early clear, then bench.py --gpus 2 (sharded q8 headline) and its reference arm
mkdir -p gpurun_out/r2_run10
timeout 200 python tools/diag.py q5 > gpurun_out/r2_run10/diag_q5.txt 2>&1
head -3 gpurun_out/r2_run10/diag_q5.txt | cut -c1-200; grep agg_ gpurun_out/r2_run10/diag_q5.txt
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 50 --warmup 5 ) > gpurun_out/r2_run10/bench_n2.json 2> gpurun_out/r2_run10/bench_n2.err
tail -12 gpurun_out/r2_run10/bench_n2.err
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 ) > gpurun_out/r2_run10/bench_ref_n2.json 2> gpurun_out/r2_run10/bench_ref_n2.err
tail -5 gpurun_out/r2_run10/bench_ref_n2.err
