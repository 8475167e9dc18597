#!/bin/bash
# round 2, run 28: verification of the code state on one GPU: smoke(), the whole -m gpu suite, the default bench line, the reference arm
O=gpurun_out/r2_run28; mkdir -p $O
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/smoke.txt 2>&1; tail -2 $O/smoke.txt
timeout 1500 python -m pytest tests -m gpu -q > $O/pytest.log 2>&1; tail -4 $O/pytest.log
timeout 600 python bench.py --impl reference > $O/bench_reference.json 2> $O/bench_reference.err; tail -c 1500 $O/bench_reference.json
timeout 900 python bench.py > $O/bench.json 2> $O/bench.err
python - <<'PY'
import json
for l in open('gpurun_out/r2_run28/bench.json'):
    if l.startswith('{'):
        d=json.loads(l); print({k:d[k] for k in ('value','ms_per_step','roofline','parity_check','gpu_launches','clocks') if k in d}); print({k:(v['value'],v['ms_per_step']) for k,v in d['e2e']['variants'].items()}); print(d['cpu_baseline']['figures'])
        for q,v in d['queries'].items(): print(q, v['ms'], v.get('ms_best'), v.get('roofline',{}).get('frac'), {k:round(x['ms'],4) for k,x in v['kernels'].items()})
PY
grep -c libflockgpu /dev/null; tail -3 $O/bench_reference.err
