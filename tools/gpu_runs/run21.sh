#!/bin/bash
# run 21 (1 GPU, lean): final verification of HEAD -- smoke, parity suite, default bench with the process-parallel CPU arm
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
( timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke21.txt 2>&1 ); tail -1 gpurun_out/smoke21.txt
( timeout 240 python -m pytest tests -m gpu -q --timeout 200 -x > gpurun_out/pytest_gpu21.txt 2>&1 ); echo "pytest rc=$?" >> gpurun_out/pytest_gpu21.txt
tail -4 gpurun_out/pytest_gpu21.txt
( timeout 150 python bench.py --e2e-steps 30 > gpurun_out/bench21.json 2> gpurun_out/bench21.err ); tail -1 gpurun_out/bench21.err | cut -c1-200
python - <<PY
import json
d=json.load(open("gpurun_out/bench21.json"))
print("value %.4g"%d["value"], "roofline", d["roofline"]["frac"], "kernel_ms", d["roofline"]["kernel_ms"], "e2e ev/s %.3g" % d["e2e"]["value"], "cpu", d.get("cpu_baseline"))
PY
