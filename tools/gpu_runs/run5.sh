#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/pytest_gpu.txt 2>&1 ); echo "pytest rc=$?" >> gpurun_out/pytest_gpu.txt
tail -12 gpurun_out/pytest_gpu.txt
( timeout 300 python tools/trace_filter.py gpurun_out/trace.txt > gpurun_out/trace_summary.txt 2>&1 ); cat gpurun_out/trace_summary.txt
for cfg in "single" "lookback"; do
  if [ "$cfg" = "lookback" ]; then export FLOCKGPU_FORCE_LOOKBACK=1; else unset FLOCKGPU_FORCE_LOOKBACK; fi
  ( timeout 600 python bench.py --steps 200 --warmup 5 --no-cpu-baseline --e2e-steps 20 > gpurun_out/bench_$cfg.json 2> gpurun_out/bench_$cfg.err )
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_$cfg.json"))
print("$cfg", "ms/step", round(d["ms_per_step"],5), "host_us", d.get("host_enqueue_us_per_step"), "roofline", d["roofline"]["frac"], "kernel_ms", d["roofline"]["kernel_ms"], "e2e ms", round(d["e2e"]["ms_per_step"],3))
PY
done
unset FLOCKGPU_FORCE_LOOKBACK
( timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_v3.csv python bench.py --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 2 > gpurun_out/ncu_bench.log 2>&1 )
( timeout 900 ncu --set full --clock-control none --import-source on -k regex:filter_compact -s 4 -c 2 -o gpurun_out/prof_filter3 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 2 > gpurun_out/ncu_full.log 2>&1 )
( timeout 900 ncu --set full --clock-control none --import-source on -k regex:agg_local32 -s 1 -c 1 -o gpurun_out/prof_agg32 python tools/nexmark_bench.py --queries q5 --reps 1 --no-e2e --no-cpu > gpurun_out/ncu_agg.log 2>&1 )
( timeout 1500 python tools/nexmark_bench.py --queries q1,q3,q5,q8 --q8-scale 0.125 --reps 10 --no-e2e --no-cpu > gpurun_out/nexmark_c.jsonl 2> gpurun_out/nexmark_c.err )
python - <<PY
import json
for l in open("gpurun_out/nexmark_c.jsonl"):
    d=json.loads(l)
    print(d["query"], "ms", round(d["device_ms_median"],4), "rows/s", "%.3g"%d["rows_per_sec"], "frac", round(d["frac_of_hbm_peak"],4))
    print("   ", {k:(v["launches"], round(v["ms"],4)) for k,v in d["kernels"].items()})
PY
head -8 gpurun_out/launches_v3.csv | cut -c1-250
