#!/bin/bash
# round 2, run 2: the full -m gpu suite on one B200 after the multi-way partition rewrite
mkdir -p gpurun_out/r2_run2
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_run2/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_run2/pytest_gpu.log
tail -15 gpurun_out/r2_run2/pytest_gpu.log
