#!/bin/bash
# round 2, run 42: the N > 1 reference arm (native q8 on the box's host cores)
mkdir -p gpurun_out/r2_run42
WORLD_SIZE=8 RANK=0 timeout 150 python bench.py --impl reference --gpus 8 --steps 5 --warmup 1 > gpurun_out/r2_run42/bench_reference_n8.json 2> gpurun_out/r2_run42/err.txt
python -c "
import json; d=json.loads(open('gpurun_out/r2_run42/bench_reference_n8.json').read().strip().splitlines()[-1]); print(d['value'], d['rows_out'], {k:(round(v['value']/1e9,3), round(v['ms_per_step'],1)) for k,v in d['cpu_baseline']['figures'].items()})"
