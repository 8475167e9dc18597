#!/bin/bash
# round 2, run 4: deferred dense aggregate (q5), finer partition grid, new bench.py (N = 1)
mkdir -p gpurun_out/r2_run4
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_run4/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_run4/pytest_gpu.log
tail -15 gpurun_out/r2_run4/pytest_gpu.log
( time timeout 900 python bench.py --steps 200 --warmup 5 ) > gpurun_out/r2_run4/bench_n1.json 2> gpurun_out/r2_run4/bench_n1.err
tail -5 gpurun_out/r2_run4/bench_n1.err
( time timeout 300 python bench.py --impl reference --steps 20 --warmup 2 ) > gpurun_out/r2_run4/bench_ref_n1.json 2> gpurun_out/r2_run4/bench_ref_n1.err
tail -4 gpurun_out/r2_run4/bench_ref_n1.err
