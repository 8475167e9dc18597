#!/bin/bash
# round 2, run 13: MAX from the scan's atomics, staged-source scatter, coalesced partition scan, multi-column Utf8 take
mkdir -p gpurun_out/r2_run13
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_nexmark.py tests/test_gpu_baseline_sizes.py -m gpu -x -q > gpurun_out/r2_run13/pytest.log 2>&1
tail -4 gpurun_out/r2_run13/pytest.log
timeout 300 python tools/diag.py q5 > gpurun_out/r2_run13/diag_q5.txt 2>&1
head -1 gpurun_out/r2_run13/diag_q5.txt | cut -c1-200; grep agg_ gpurun_out/r2_run13/diag_q5.txt
timeout 300 python tools/diag.py partition > gpurun_out/r2_run13/diag_partition.txt 2>&1
cat gpurun_out/r2_run13/diag_partition.txt
timeout 300 python tools/nexmark_bench.py --queries q8,q3 --q8-scale 0.125 --reps 20 --no-cpu --no-e2e > gpurun_out/r2_run13/nexmark.jsonl 2> gpurun_out/r2_run13/nexmark.err
python -c "
import json
for l in open('gpurun_out/r2_run13/nexmark.jsonl'):
    d=json.loads(l); print(d['query'], round(d['device_ms_median'],4), {k:(v['launches'], round(v['ms'],4)) for k,v in d['kernels'].items()})
"
