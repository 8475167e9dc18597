#!/bin/bash
# round 2, run 18: NDJSON parse, C ABI program from C, full suite
mkdir -p gpurun_out/r2_run18
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2_run18/pytest.log 2>&1
tail -30 gpurun_out/r2_run18/pytest.log
