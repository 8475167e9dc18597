#!/bin/bash
# round 2, run 19: validity bitmaps (NULL semantics against pyarrow), full suite, then a short bench for regressions
mkdir -p gpurun_out/r2_run19
timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -q -x -k "nulls or unknown_types" > gpurun_out/r2_run19/nulls.log 2>&1
tail -40 gpurun_out/r2_run19/nulls.log
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2_run19/pytest.log 2>&1
tail -30 gpurun_out/r2_run19/pytest.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/r2_run19/bench.json 2> gpurun_out/r2_run19/bench.err
tail -c 3000 gpurun_out/r2_run19/bench.json
