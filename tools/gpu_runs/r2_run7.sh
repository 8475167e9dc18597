#!/bin/bash
# round 2, run 7: block allocator, device-side dense range, fused argmax; q5 per-repetition times; pageable feed threads
mkdir -p gpurun_out/r2_run7
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2_run7/pytest.log 2>&1
tail -8 gpurun_out/r2_run7/pytest.log
FLOCKGPU_HOST_TRACE=1 timeout 300 python tools/diag.py q5 > gpurun_out/r2_run7/diag_q5.txt 2>&1
cat gpurun_out/r2_run7/diag_q5.txt
timeout 300 python tools/diag.py feed > gpurun_out/r2_run7/diag_feed.txt 2>&1
cat gpurun_out/r2_run7/diag_feed.txt
