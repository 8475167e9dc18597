#!/bin/bash
# round 2, run 12: dense set-up variants for q5; scatter kernel with batched loads
mkdir -p gpurun_out/r2_run12
for v in 0 1 2; do
  echo "== FLOCKGPU_DENSE_VARIANT=$v"
  FLOCKGPU_DENSE_VARIANT=$v timeout 200 python tools/diag.py q5 > gpurun_out/r2_run12/diag_q5_v$v.txt 2>&1
  head -1 gpurun_out/r2_run12/diag_q5_v$v.txt | cut -c1-160; grep agg_ gpurun_out/r2_run12/diag_q5_v$v.txt
done
timeout 300 python tools/diag.py partition > gpurun_out/r2_run12/diag_partition.txt 2>&1
cat gpurun_out/r2_run12/diag_partition.txt
timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -x -q -k "partition or prefix" > gpurun_out/r2_run12/pytest.log 2>&1
tail -3 gpurun_out/r2_run12/pytest.log
