#!/bin/bash
# round 2, run 17: SortExec / ROW_NUMBER / LIMIT (q6), IPC frames, window assembly -- full -m gpu suite
mkdir -p gpurun_out/r2_run17
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2_run17/pytest.log 2>&1
tail -25 gpurun_out/r2_run17/pytest.log
