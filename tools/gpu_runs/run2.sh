#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/pytest_gpu.txt 2>&1 ); echo "pytest rc=$?" >> gpurun_out/pytest_gpu.txt
for it in 16 32 64; do
  ( FLOCKGPU_FILTER_ITEMS=$it timeout 600 python bench.py --steps 200 --warmup 5 --no-cpu-baseline --e2e-steps 20 > gpurun_out/bench_items$it.json 2> gpurun_out/bench_items$it.err )
done
( timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv python bench.py --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 2 > gpurun_out/ncu_bench.log 2>&1 )
( timeout 900 ncu --set full --clock-control none --import-source on -k regex:filter_compact -s 4 -c 3 -o gpurun_out/prof_filter python bench.py --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 2 > gpurun_out/ncu_full.log 2>&1 )
tail -4 gpurun_out/pytest_gpu.txt
for it in 16 32 64; do python - <<PY
import json
d=json.load(open("gpurun_out/bench_items$it.json"))
print($it, "value", d["value"], "ms/step", d["ms_per_step"], "host_us", d.get("host_enqueue_us_per_step"), "roofline", d["roofline"], "e2e", d["e2e"]["value"])
PY
done
head -30 gpurun_out/launches.csv | cut -c1-300
