#!/bin/bash
# round 2, run 1: BASELINE-size parity tests on the round-1 library + current per-query numbers
mkdir -p gpurun_out/r2_run1
timeout 1500 python -m pytest tests/test_gpu_baseline_sizes.py -m gpu -x -q > gpurun_out/r2_run1/pytest_baseline.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_run1/pytest_baseline.log
timeout 600 python tools/nexmark_bench.py --queries q1,q2,q3,q5 --reps 20 --no-cpu --no-e2e > gpurun_out/r2_run1/nexmark.jsonl 2> gpurun_out/r2_run1/nexmark.err
timeout 600 python tools/nexmark_bench.py --queries q8 --q8-scale 0.125 --reps 20 --no-cpu --no-e2e >> gpurun_out/r2_run1/nexmark.jsonl 2>> gpurun_out/r2_run1/nexmark.err
tail -5 gpurun_out/r2_run1/pytest_baseline.log
