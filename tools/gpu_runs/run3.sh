#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/pytest_gpu.txt 2>&1 ); echo "pytest rc=$?" >> gpurun_out/pytest_gpu.txt
for it in 32 64; do
  ( FLOCKGPU_FILTER_ITEMS=$it timeout 600 python bench.py --steps 200 --warmup 5 --no-cpu-baseline --e2e-steps 20 > gpurun_out/bench_items$it.json 2> gpurun_out/bench_items$it.err )
done
( timeout 900 ncu --set full --clock-control none --import-source on -k regex:filter_compact -s 4 -c 2 -o gpurun_out/prof_filter2 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 2 > gpurun_out/ncu_full.log 2>&1 )
( timeout 1500 python tools/nexmark_bench.py --queries q1,q2,q3,q5,q8 --q8-scale 0.125 --reps 10 > gpurun_out/nexmark_all.jsonl 2> gpurun_out/nexmark_all.err )
tail -3 gpurun_out/pytest_gpu.txt
for it in 32 64; do python - <<PY
import json
d=json.load(open("gpurun_out/bench_items$it.json"))
print($it, "value", d["value"], "ms/step", d["ms_per_step"], "host_us", d.get("host_enqueue_us_per_step"), "roofline", d["roofline"]["frac"], d["roofline"]["kernel_ms"], "e2e", d["e2e"]["value"])
PY
done
python - <<PY
import json
for l in open("gpurun_out/nexmark_all.jsonl"):
    d=json.loads(l)
    print(d["query"], "ms", round(d["device_ms_median"],4), "rows/s", "%.3g"%d["rows_per_sec"], "frac", round(d["frac_of_hbm_peak"],4), "e2e_ms", d["e2e_ms"], "cpu", d.get("cpu_oracle"))
    print("   ", {k:(v["launches"], round(v["ms"],4)) for k,v in d["kernels"].items()})
PY
tail -5 gpurun_out/nexmark_all.err
