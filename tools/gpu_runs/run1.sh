#!/bin/bash
# first GPU run: tests, smoke, microbench, bench, ncu
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
( timeout 120 ./tools/microbench > gpurun_out/microbench.txt 2>&1 ) 
( timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.txt 2>&1 ); echo "smoke rc=$?" >> gpurun_out/smoke.txt
( timeout 1500 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/pytest_gpu.txt 2>&1 ); echo "pytest rc=$?" >> gpurun_out/pytest_gpu.txt
( timeout 600 python bench.py --steps 200 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err ); echo "bench rc=$?" >> gpurun_out/bench.err
tail -5 gpurun_out/pytest_gpu.txt; cat gpurun_out/smoke.txt | tail -3; cat gpurun_out/bench.json | cut -c1-1500; tail -5 gpurun_out/bench.err
