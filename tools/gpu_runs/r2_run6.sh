#!/bin/bash
mkdir -p gpurun_out/r2_run6
FLOCKGPU_HOST_TRACE=1 timeout 300 python tools/diag.py q5 > gpurun_out/r2_run6/diag_q5.txt 2>&1
cat gpurun_out/r2_run6/diag_q5.txt
