#!/bin/bash
# round 2, run 8: q5 after the register-resident argmax, feed staging sweep, ncu of the q5 kernels
mkdir -p gpurun_out/r2_run8
timeout 600 python -m pytest tests/test_gpu_ops.py tests/test_gpu_nexmark.py -m gpu -x -q -k "dense or q5 or nexmark or zero_copy" > gpurun_out/r2_run8/pytest.log 2>&1
tail -4 gpurun_out/r2_run8/pytest.log
timeout 300 python tools/diag.py q5 > gpurun_out/r2_run8/diag_q5.txt 2>&1
head -3 gpurun_out/r2_run8/diag_q5.txt | cut -c1-200; grep agg_ gpurun_out/r2_run8/diag_q5.txt
timeout 300 python tools/diag.py feed > gpurun_out/r2_run8/diag_feed.txt 2>&1
cat gpurun_out/r2_run8/diag_feed.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_run8/q5_launches.csv python tools/prof_query.py q5 3 > gpurun_out/r2_run8/ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:agg_hist32 -s 2 -c 1 -o gpurun_out/r2_run8/q5_hist python tools/prof_query.py q5 3 > gpurun_out/r2_run8/ncu_hist.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:dense_argmax -s 2 -c 1 -o gpurun_out/r2_run8/q5_argmax python tools/prof_query.py q5 3 > gpurun_out/r2_run8/ncu_argmax.log 2>&1
ls -la gpurun_out/r2_run8/
