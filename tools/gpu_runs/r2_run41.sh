#!/bin/bash
# round 2, run 41: the IPC / NULL round-trip tests after the frame validation moved in front of the context
mkdir -p gpurun_out/r2_run41
timeout 200 python -m pytest tests/test_gpu_ops.py tests/test_gpu_nexmark.py -m gpu -q -x -k "ipc or nulls_roundtrip or frames" > gpurun_out/r2_run41/ipc.log 2>&1; tail -3 gpurun_out/r2_run41/ipc.log
