#!/bin/bash
# round 2, run 24: relaxed window vote in the dense histogram kernel (A/B against the strict per-step vote), full suite, N = 1 bench
mkdir -p gpurun_out/r2_run24
timeout 900 python -m pytest tests -m gpu -q -x -k "aggregate or q5 or q8 or dense or distinct or baseline" > gpurun_out/r2_run24/focus.log 2>&1
tail -8 gpurun_out/r2_run24/focus.log
FLOCKGPU_HIST_STRICT=1 timeout 300 python tools/diag.py q5 > gpurun_out/r2_run24/diag_q5_strict.txt 2>&1
timeout 300 python tools/diag.py q5 > gpurun_out/r2_run24/diag_q5_relaxed.txt 2>&1
FLOCKGPU_HIST_STRICT=1 timeout 300 python tools/diag.py q5 > gpurun_out/r2_run24/diag_q5_strict2.txt 2>&1
timeout 300 python tools/diag.py q5 > gpurun_out/r2_run24/diag_q5_relaxed2.txt 2>&1
for f in strict relaxed strict2 relaxed2; do echo "== $f"; grep -E "^flush|^warm|agg_hist" gpurun_out/r2_run24/diag_q5_$f.txt | cut -c1-260; done
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2_run24/pytest.log 2>&1
tail -5 gpurun_out/r2_run24/pytest.log
timeout 600 python bench.py > gpurun_out/r2_run24/bench.json 2> gpurun_out/r2_run24/bench.err
python - <<'PY'
import json
for l in open('gpurun_out/r2_run24/bench.json'):
    if l.startswith('{'):
        d=json.loads(l); print({k:d[k] for k in ('value','ms_per_step','e2e','roofline','parity_check') if k in d})
        for q,v in d['queries'].items(): print(q, v['ms'], v.get('ms_best'), v.get('roofline',{}).get('frac'), {k:round(x['ms'],4) for k,x in v['kernels'].items()})
PY
