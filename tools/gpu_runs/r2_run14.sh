#!/bin/bash
# round 2, run 14: ncu of the partition kernels on 2.5 M (p_id, name) rows
mkdir -p gpurun_out/r2_run14
timeout 900 ncu --set full --clock-control none --import-source on -k regex:partition_scatter -s 3 -c 1 -o gpurun_out/r2_run14/scatter python tools/diag.py partition > gpurun_out/r2_run14/ncu_scatter.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:partition_count -s 3 -c 1 -o gpurun_out/r2_run14/count python tools/diag.py partition > gpurun_out/r2_run14/ncu_count.log 2>&1
ls -la gpurun_out/r2_run14
