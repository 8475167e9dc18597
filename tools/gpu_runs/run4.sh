#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/pytest_gpu.txt 2>&1 ); echo "pytest rc=$?" >> gpurun_out/pytest_gpu.txt
tail -15 gpurun_out/pytest_gpu.txt
( timeout 300 python tools/trace_filter.py gpurun_out/trace.txt > gpurun_out/trace_summary.txt 2>&1 ); cat gpurun_out/trace_summary.txt
for cfg in "32 100" "1 0" "32 0" "1 100"; do
  set -- $cfg
  ( FLOCKGPU_LB_STRIDE=$1 FLOCKGPU_LB_SLEEP=$2 timeout 600 python bench.py --steps 200 --warmup 5 --no-cpu-baseline --e2e-steps 20 > gpurun_out/bench_lb_$1_$2.json 2> gpurun_out/bench_lb_$1_$2.err )
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_lb_$1_$2.json"))
print("stride $1 sleep $2", "ms/step", round(d["ms_per_step"],5), "host_us", d.get("host_enqueue_us_per_step"), "roofline", d["roofline"]["frac"], "kernel_ms", d["roofline"]["kernel_ms"], "e2e ms", round(d["e2e"]["ms_per_step"],3))
PY
done
( timeout 1500 python tools/nexmark_bench.py --queries q3,q5,q8 --q8-scale 0.125 --reps 10 --no-e2e > gpurun_out/nexmark_b.jsonl 2> gpurun_out/nexmark_b.err )
python - <<PY
import json
for l in open("gpurun_out/nexmark_b.jsonl"):
    d=json.loads(l)
    print(d["query"], "ms", round(d["device_ms_median"],4), "rows/s", "%.3g"%d["rows_per_sec"], "frac", round(d["frac_of_hbm_peak"],4))
    print("   ", {k:(v["launches"], round(v["ms"],4)) for k,v in d["kernels"].items()})
PY
tail -5 gpurun_out/nexmark_b.err
