#!/bin/bash
# round 2, run 35: the default bench line of the final code state (roofline from the back-to-back region, event pairs beside it) and the reference arm
O=gpurun_out/r2_run35; mkdir -p $O
timeout 900 python bench.py > $O/bench.json 2> $O/bench.err
python - <<'PY'
import json
for l in open('gpurun_out/r2_run35/bench.json'):
    if l.startswith('{'):
        d=json.loads(l); print({k:d[k] for k in ('value','ms_per_step','ms_per_step_instrumented','roofline','parity_check','gpu_launches','clocks') if k in d}); print({k:(v['value'],v['ms_per_step']) for k,v in d['e2e']['variants'].items()}); print(d['cpu_baseline']['figures'])
        for q,v in d['queries'].items(): print(q, v['ms'], v.get('ms_best'), v.get('roofline',{}).get('frac'))
PY
timeout 600 python bench.py --impl reference --steps 50 --warmup 3 > $O/bench_reference.json 2> $O/bench_reference.err; tail -c 700 $O/bench_reference.json
