#!/bin/bash
# round 2, run 5: q5 per-repetition times, pageable feed vs staging threads, dense-aggregate tests
mkdir -p gpurun_out/r2_run5
timeout 600 python -m pytest tests/test_gpu_ops.py tests/test_gpu_baseline_sizes.py -m gpu -x -q -k "dense or q5 or q8 or aggregate" > gpurun_out/r2_run5/pytest.log 2>&1
tail -5 gpurun_out/r2_run5/pytest.log
timeout 300 python tools/diag.py q5 > gpurun_out/r2_run5/diag_q5.txt 2>&1
cat gpurun_out/r2_run5/diag_q5.txt
timeout 300 python tools/diag.py feed > gpurun_out/r2_run5/diag_feed.txt 2>&1
cat gpurun_out/r2_run5/diag_feed.txt
