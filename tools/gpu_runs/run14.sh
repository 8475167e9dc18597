#!/bin/bash
# run 14 (1 GPU): one-row-build join, 32-bit partials, unrolled global aggregate; driver-style bench twice + smoke + reference arm
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
( timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke14.txt 2>&1 ); tail -2 gpurun_out/smoke14.txt
( timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/pytest_gpu14.txt 2>&1 ); echo "pytest rc=$?" >> gpurun_out/pytest_gpu14.txt
tail -12 gpurun_out/pytest_gpu14.txt
for i in a b; do
  ( timeout 900 python bench.py > gpurun_out/bench14_$i.json 2> gpurun_out/bench14_$i.err ); tail -1 gpurun_out/bench14_$i.err | cut -c1-200
  python - <<PY
import json
d=json.load(open("gpurun_out/bench14_$i.json"))
print("$i value %.4g"%d["value"], "ms/step", round(d["ms_per_step"],5), "roofline", d["roofline"]["frac"], "kernel_ms", d["roofline"]["kernel_ms"], "e2e ev/s %.3g" % d["e2e"]["value"], d["e2e"]["ms_per_step"], d["e2e"].get("host_ms_per_step"), "cpu", d.get("cpu_baseline",{}).get("value"))
PY
done
( timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench14_ref.json 2> gpurun_out/bench14_ref.err ); cut -c1-300 gpurun_out/bench14_ref.json
( timeout 1500 python tools/nexmark_bench.py --queries q1,q2,q3,q5,q8 --q8-scale 0.125 --reps 10 --no-cpu --no-e2e > gpurun_out/nexmark14.jsonl 2> gpurun_out/nexmark14.err )
python - <<PY
import json
for l in open("gpurun_out/nexmark14.jsonl"):
    if not l.startswith("{"): continue
    d=json.loads(l)
    print(d["query"], "ms", round(d["device_ms_median"],4), "rows/s", "%.3g"%d["rows_per_sec"], "frac", round(d["frac_of_hbm_peak"],4))
    print("   ", {k:(v["launches"], round(v["ms"],4)) for k,v in d["kernels"].items()})
PY
tail -3 gpurun_out/nexmark14.err
