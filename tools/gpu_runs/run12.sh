#!/bin/bash
# run 12 (1 GPU): relaxed arrival counter prefix, lean emit kernel, adaptive tiles, Final without level 1
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/pytest_gpu12.txt 2>&1 ); echo "pytest rc=$?" >> gpurun_out/pytest_gpu12.txt
tail -8 gpurun_out/pytest_gpu12.txt
( timeout 300 python tools/trace_filter.py gpurun_out/trace12.txt > gpurun_out/trace12_summary.txt 2>&1 ); cat gpurun_out/trace12_summary.txt
for cfg in "zc" "copy"; do
  if [ "$cfg" = "copy" ]; then X="--e2e-copy"; else X=""; fi
  ( timeout 600 python bench.py --steps 200 --warmup 5 --no-cpu-baseline --e2e-steps 50 $X > gpurun_out/bench12_$cfg.json 2> gpurun_out/bench12_$cfg.err )
  python - <<PY
import json
d=json.load(open("gpurun_out/bench12_$cfg.json"))
print("$cfg", "value %.4g"%d["value"], "ms/step", round(d["ms_per_step"],5), "host_us", d.get("host_enqueue_us_per_step"), "roofline", d["roofline"]["frac"], "kernel_ms", d["roofline"]["kernel_ms"], "e2e ms", round(d["e2e"]["ms_per_step"],3), "e2e ev/s %.3g" % d["e2e"]["value"], d["e2e"].get("host_ms_per_step"))
PY
done
( timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches12.csv python bench.py --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 2 > gpurun_out/ncu_bench12.log 2>&1 )
( timeout 900 ncu --set full --clock-control none --import-source on -k regex:filter_compact -s 4 -c 2 -o gpurun_out/prof_filter12 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 2 > gpurun_out/ncu_full12.log 2>&1 )
( timeout 1500 python tools/nexmark_bench.py --queries q1,q2,q3,q5,q8 --q8-scale 0.125 --reps 10 --no-cpu > gpurun_out/nexmark12.jsonl 2> gpurun_out/nexmark12.err )
python - <<PY
import json
for l in open("gpurun_out/nexmark12.jsonl"):
    if not l.startswith("{"): continue
    d=json.loads(l)
    print(d["query"], "ms", round(d["device_ms_median"],4), "rows/s", "%.3g"%d["rows_per_sec"], "frac", round(d["frac_of_hbm_peak"],4), "e2e_ms", d["e2e_ms"])
    print("   ", {k:(v["launches"], round(v["ms"],4)) for k,v in d["kernels"].items()})
PY
tail -3 gpurun_out/nexmark12.err
( timeout 900 ncu --set full --clock-control none --import-source on -k regex:"agg_hist32|agg_emit_kernel|join_count_scan|agg_insert_kernel" -c 4 -o gpurun_out/prof_q5_12 python tools/nexmark_bench.py --queries q5 --reps 1 --no-e2e --no-cpu > gpurun_out/ncu_q5_12.log 2>&1 ); tail -2 gpurun_out/ncu_q5_12.log
