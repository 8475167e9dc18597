#!/bin/bash
# round 2, run 9: sampling + clearing fused into the scan kernel; staging pool
mkdir -p gpurun_out/r2_run9
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_nexmark.py tests/test_gpu_baseline_sizes.py -m gpu -x -q > gpurun_out/r2_run9/pytest.log 2>&1
tail -4 gpurun_out/r2_run9/pytest.log
timeout 300 python tools/diag.py q5 > gpurun_out/r2_run9/diag_q5.txt 2>&1
head -3 gpurun_out/r2_run9/diag_q5.txt | cut -c1-200; grep agg_ gpurun_out/r2_run9/diag_q5.txt
timeout 300 python tools/diag.py feed > gpurun_out/r2_run9/diag_feed.txt 2>&1
cat gpurun_out/r2_run9/diag_feed.txt
