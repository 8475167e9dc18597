#!/bin/bash
# run 20 (1 GPU, lean): validates the kernel materialisation of host-chunked columns and the pred_i32.h refactor
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_nexmark.py tests/test_gpu_ops.py -m gpu -q --timeout 300 -k "zero_copy or filter or divisib or golden or matches_oracle or prefix" > gpurun_out/pytest_gpu20.txt 2>&1 ); echo "pytest rc=$?" >> gpurun_out/pytest_gpu20.txt
tail -6 gpurun_out/pytest_gpu20.txt
( timeout 600 python tools/nexmark_bench.py --queries q2,q5,q3 --reps 5 --no-cpu > gpurun_out/nexmark20.jsonl 2> gpurun_out/nexmark20.err )
python - <<PY
import json
for l in open("gpurun_out/nexmark20.jsonl"):
    if not l.startswith("{"): continue
    d=json.loads(l)
    print(d["query"], "ms", round(d["device_ms_median"],4), "e2e_ms", d["e2e_ms"], "e2e rows/s", d["e2e_rows_per_sec"])
PY
tail -2 gpurun_out/nexmark20.err
