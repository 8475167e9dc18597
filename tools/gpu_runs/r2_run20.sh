#!/bin/bash
# round 2, run 20: NULL semantics again (accumulator sharing), uniqueness probe, fused count+scan, 2-GPU exchange with validity, 2-GPU bench
mkdir -p gpurun_out/r2_run20
timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -q -k "nulls or unknown_types or unique_probe or partition" > gpurun_out/r2_run20/focus.log 2>&1
tail -30 gpurun_out/r2_run20/focus.log
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q -x > gpurun_out/r2_run20/multi.log 2>&1
tail -60 gpurun_out/r2_run20/multi.log
timeout 1500 python -m pytest tests -m gpu -q --deselect tests/test_gpu_multi.py > gpurun_out/r2_run20/pytest.log 2>&1
tail -15 gpurun_out/r2_run20/pytest.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r2_run20/bench_n2.json 2> gpurun_out/r2_run20/bench_n2.err
tail -c 2500 gpurun_out/r2_run20/bench_n2.json; tail -5 gpurun_out/r2_run20/bench_n2.err
