#!/bin/bash
# round 2, run 43: the default bench line with the per-query CPU figures in cpu_baseline
mkdir -p gpurun_out/r2_run43
timeout 125 python bench.py > gpurun_out/r2_run43/bench.json 2> gpurun_out/r2_run43/bench.err
python -c "
import json
d=json.loads([l for l in open('gpurun_out/r2_run43/bench.json') if l.startswith('{')][-1]); print(d['value'], d['e2e']['value'], d['roofline']['frac'], d['parity_check']); print(json.dumps(d['cpu_baseline'].get('queries'))[:900]); print({k:v['value'] for k,v in d['cpu_baseline']['figures'].items()})"
tail -2 gpurun_out/r2_run43/bench.err
