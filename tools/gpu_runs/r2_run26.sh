#!/bin/bash
# round 2, run 26: join table with the key in the slot (4-byte keys); full suite; default bench
O=gpurun_out/r2_run26; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -x -k "join or q3 or q8 or q4 or q7 or nexmark or golden" > $O/focus.log 2>&1
tail -5 $O/focus.log
timeout 1500 python -m pytest tests -m gpu -q > $O/pytest.log 2>&1
tail -5 $O/pytest.log
timeout 600 python bench.py > $O/bench.json 2> $O/bench.err
python - <<'PY'
import json
for l in open('gpurun_out/r2_run26/bench.json'):
    if l.startswith('{'):
        d=json.loads(l); print({k:d[k] for k in ('value','ms_per_step','roofline','parity_check','gpu_launches') if k in d}); print({k:(v['value'],v['ms_per_step']) for k,v in d['e2e']['variants'].items()}); print(d['cpu_baseline']['figures'])
        for q,v in d['queries'].items(): print(q, v['ms'], v.get('ms_best'), v.get('roofline',{}).get('frac'), {k:round(x['ms'],4) for k,x in v['kernels'].items()})
PY
