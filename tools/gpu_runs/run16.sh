#!/bin/bash
# run 16 (1 GPU): balanced survivor write (A/B), emit with adaptive tiles, large-table tests
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/pytest_gpu16.txt 2>&1 ); echo "pytest rc=$?" >> gpurun_out/pytest_gpu16.txt
tail -12 gpurun_out/pytest_gpu16.txt
( timeout 300 python tools/trace_filter.py gpurun_out/trace16.txt > gpurun_out/trace16_summary.txt 2>&1 ); cat gpurun_out/trace16_summary.txt
for cfg in balanced plain; do
  if [ "$cfg" = "plain" ]; then export FLOCKGPU_NO_BALANCED=1; else unset FLOCKGPU_NO_BALANCED; fi
  ( timeout 600 python bench.py --steps 200 --warmup 5 --no-cpu-baseline --e2e-steps 20 > gpurun_out/bench16_$cfg.json 2> gpurun_out/bench16_$cfg.err )
  python - <<PY
import json
d=json.load(open("gpurun_out/bench16_$cfg.json"))
print("$cfg", "value %.4g"%d["value"], "ms/step", round(d["ms_per_step"],5), "roofline", d["roofline"]["frac"], "kernel_ms", d["roofline"]["kernel_ms"], "e2e ev/s %.3g" % d["e2e"]["value"])
PY
done
unset FLOCKGPU_NO_BALANCED
( timeout 900 ncu --set full --clock-control none --import-source on -k regex:filter_compact -s 4 -c 1 -o gpurun_out/prof_filter16 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 2 > gpurun_out/ncu_full16.log 2>&1 )
( timeout 1500 python tools/nexmark_bench.py --queries q2,q3,q5,q8 --q8-scale 0.125 --reps 10 --no-cpu --no-e2e > gpurun_out/nexmark16.jsonl 2> gpurun_out/nexmark16.err )
python - <<PY
import json
for l in open("gpurun_out/nexmark16.jsonl"):
    if not l.startswith("{"): continue
    d=json.loads(l)
    print(d["query"], "ms", round(d["device_ms_median"],4), "rows/s", "%.3g"%d["rows_per_sec"], "frac", round(d["frac_of_hbm_peak"],4))
    print("   ", {k:(v["launches"], round(v["ms"],4)) for k,v in d["kernels"].items()})
PY
