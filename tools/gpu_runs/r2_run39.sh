#!/bin/bash
# round 2, run 39: the default bench line and the reference arm of the final code state (clock sampling every 20 ms during the e2e legs)
O=gpurun_out/r2_run39; mkdir -p $O
timeout 900 python bench.py > $O/bench.json 2> $O/bench.err
python - <<'PY'
import json
for l in open('gpurun_out/r2_run39/bench.json'):
    if l.startswith('{'):
        d=json.loads(l); print({k:d[k] for k in ('value','ms_per_step','roofline','parity_check','gpu_launches','clocks') if k in d}); print({k:(v['value'],v['ms_per_step'],v['host_ms_per_step']) for k,v in d['e2e']['variants'].items()}); print(d['cpu_baseline']['figures'])
        for q,v in d['queries'].items(): print(q, v['ms'], v.get('ms_best'), v.get('roofline',{}).get('frac'))
PY
timeout 600 python bench.py --impl reference --steps 50 --warmup 3 > $O/bench_reference.json 2> $O/bench_reference.err; python -c "
import json
d=json.loads(open('gpurun_out/r2_run39/bench_reference.json').read().strip().splitlines()[-1]); print(d['value'], {k:v['value'] for k,v in d['cpu_baseline']['figures'].items()})"
