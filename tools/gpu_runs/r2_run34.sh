#!/bin/bash
# round 2, run 34: verification of the final code state on one GPU (bench with the two timed regions and the Acero figure)
O=gpurun_out/r2_run34; mkdir -p $O
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/smoke.txt 2>&1; tail -1 $O/smoke.txt
timeout 1500 python -m pytest tests -m gpu -q > $O/pytest.log 2>&1; tail -3 $O/pytest.log
timeout 900 python bench.py > $O/bench.json 2> $O/bench.err
python - <<'PY'
import json
for l in open('gpurun_out/r2_run34/bench.json'):
    if l.startswith('{'):
        d=json.loads(l); print({k:d[k] for k in ('value','ms_per_step','ms_per_step_instrumented','roofline','parity_check','gpu_launches') if k in d}); print({k:(v['value'],v['ms_per_step']) for k,v in d['e2e']['variants'].items()}); print(d['cpu_baseline']['figures'])
        for q,v in d['queries'].items(): print(q, v['ms'], v.get('ms_best'), v.get('roofline',{}).get('frac'), {k:round(x['ms'],4) for k,x in v['kernels'].items()})
PY
tail -2 $O/bench.err
