#!/bin/bash
# run 17 (1 GPU): fused take() (one fixed-width launch, batched Utf8 totals), Utf8 copy through shared staging, PCIe read microbench
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/pytest_gpu17.txt 2>&1 ); echo "pytest rc=$?" >> gpurun_out/pytest_gpu17.txt
tail -12 gpurun_out/pytest_gpu17.txt
( timeout 300 ./build/microbench pcie > gpurun_out/microbench_pcie.txt 2>&1 ); cat gpurun_out/microbench_pcie.txt
( timeout 600 python bench.py --steps 200 --warmup 5 --no-cpu-baseline --e2e-steps 20 > gpurun_out/bench17.json 2> gpurun_out/bench17.err )
python - <<PY
import json
d=json.load(open("gpurun_out/bench17.json"))
print("value %.4g"%d["value"], "ms/step", round(d["ms_per_step"],5), "roofline", d["roofline"]["frac"], "kernel_ms", d["roofline"]["kernel_ms"], "e2e ev/s %.3g" % d["e2e"]["value"])
PY
( timeout 1500 python tools/nexmark_bench.py --queries q1,q2,q3,q5,q8 --q8-scale 0.125 --reps 10 --no-cpu --no-e2e > gpurun_out/nexmark17.jsonl 2> gpurun_out/nexmark17.err )
python - <<PY
import json
for l in open("gpurun_out/nexmark17.jsonl"):
    if not l.startswith("{"): continue
    d=json.loads(l)
    print(d["query"], "ms", round(d["device_ms_median"],4), "rows/s", "%.3g"%d["rows_per_sec"], "frac", round(d["frac_of_hbm_peak"],4))
    print("   ", {k:(v["launches"], round(v["ms"],4)) for k,v in d["kernels"].items()})
PY
tail -3 gpurun_out/nexmark17.err
