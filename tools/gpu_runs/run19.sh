#!/bin/bash
# run 19 (1 GPU): kernel materialisation of host-chunked columns, per-query e2e through the page-locked zero-copy feed
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/pytest_gpu19.txt 2>&1 ); echo "pytest rc=$?" >> gpurun_out/pytest_gpu19.txt
tail -6 gpurun_out/pytest_gpu19.txt
( timeout 1500 python tools/nexmark_bench.py --queries q1,q2,q3,q5,q8 --q8-scale 0.125 --reps 10 --no-cpu > gpurun_out/nexmark19.jsonl 2> gpurun_out/nexmark19.err )
python - <<PY
import json
for l in open("gpurun_out/nexmark19.jsonl"):
    if not l.startswith("{"): continue
    d=json.loads(l)
    print(d["query"], "ms", round(d["device_ms_median"],4), "rows/s", "%.3g"%d["rows_per_sec"], "frac", round(d["frac_of_hbm_peak"],4), "e2e_ms", d["e2e_ms"], "e2e rows/s", d["e2e_rows_per_sec"])
PY
tail -3 gpurun_out/nexmark19.err
( timeout 600 python bench.py --steps 200 --warmup 5 --no-cpu-baseline --e2e-steps 30 > gpurun_out/bench19.json 2> gpurun_out/bench19.err )
python - <<PY
import json
d=json.load(open("gpurun_out/bench19.json"))
print("value %.4g"%d["value"], "roofline", d["roofline"]["frac"], "e2e ev/s %.3g" % d["e2e"]["value"])
PY
