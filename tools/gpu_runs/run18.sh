#!/bin/bash
# run 18 (1 GPU): pinned slab arena (e2e), final-state verification: smoke, parity suite, default bench, reference arm, all queries, ncu captures
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
( timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke18.txt 2>&1 ); tail -2 gpurun_out/smoke18.txt
( timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/pytest_gpu18.txt 2>&1 ); echo "pytest rc=$?" >> gpurun_out/pytest_gpu18.txt
tail -6 gpurun_out/pytest_gpu18.txt
for cfg in zc copy; do
  if [ "$cfg" = "copy" ]; then X="--e2e-copy --no-cpu-baseline"; else X=""; fi
  ( timeout 900 python bench.py $X > gpurun_out/bench18_$cfg.json 2> gpurun_out/bench18_$cfg.err ); tail -1 gpurun_out/bench18_$cfg.err | cut -c1-200
  python - <<PY
import json
d=json.load(open("gpurun_out/bench18_$cfg.json"))
print("$cfg value %.4g"%d["value"], "ms/step", round(d["ms_per_step"],5), "roofline", d["roofline"]["frac"], "kernel_ms", d["roofline"]["kernel_ms"], "e2e ev/s %.3g" % d["e2e"]["value"], round(d["e2e"]["ms_per_step"],4), d["e2e"].get("host_ms_per_step"), {k:round(v["ms"]/v["launches"],4) for k,v in d["e2e"]["kernels"].items()}, "cpu", d.get("cpu_baseline",{}).get("value"))
PY
done
( timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench18_ref.json 2> gpurun_out/bench18_ref.err ); cut -c1-200 gpurun_out/bench18_ref.json
( timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches18.csv python bench.py --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 2 > gpurun_out/ncu_bench18.log 2>&1 )
( timeout 900 ncu --set full --clock-control none --import-source on -k regex:filter_compact -s 4 -c 1 -o gpurun_out/prof_filter18 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 2 > gpurun_out/ncu_full18.log 2>&1 )
( timeout 1500 python tools/nexmark_bench.py --queries q1,q2,q3,q5,q8 --q8-scale 0.125 --reps 10 --no-cpu > gpurun_out/nexmark18.jsonl 2> gpurun_out/nexmark18.err )
python - <<PY
import json
for l in open("gpurun_out/nexmark18.jsonl"):
    if not l.startswith("{"): continue
    d=json.loads(l)
    print(d["query"], "ms", round(d["device_ms_median"],4), "rows/s", "%.3g"%d["rows_per_sec"], "frac", round(d["frac_of_hbm_peak"],4), "e2e_ms", d["e2e_ms"])
PY
( timeout 900 ncu --set full --clock-control none --import-source on -k regex:"agg_hist32|agg_emit_kernel|agg_insert_kernel|join_one" -c 4 -o gpurun_out/prof_q5_18 python tools/nexmark_bench.py --queries q5 --reps 1 --no-e2e --no-cpu > gpurun_out/ncu_q5_18.log 2>&1 ); tail -1 gpurun_out/ncu_q5_18.log
