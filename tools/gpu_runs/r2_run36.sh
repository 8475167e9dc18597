#!/bin/bash
# round 2, run 36: pageable feed with non-temporal staging stores vs memcpy (A/B, twice), then the import / e2e tests and the e2e legs of the bench
O=gpurun_out/r2_run36; mkdir -p $O
for i in 1 2; do
  DIAG_STREAM=0 DIAG_THREADS=4,8,16 timeout 300 python tools/diag.py feed > $O/feed_memcpy$i.txt 2>&1; cat $O/feed_memcpy$i.txt
  DIAG_STREAM=1 DIAG_THREADS=4,8,16 timeout 300 python tools/diag.py feed > $O/feed_stream$i.txt 2>&1; cat $O/feed_stream$i.txt
done
timeout 900 python -m pytest tests -m gpu -q -x -k "import or roundtrip or feed or q2_full or cabi or zero_copy or nexmark_matches" > $O/focus.log 2>&1; tail -3 $O/focus.log
timeout 600 python bench.py --no-queries --no-cpu-baseline --steps 50 > $O/bench.json 2> $O/bench.err
python - <<'PY'
import json
for l in open('gpurun_out/r2_run36/bench.json'):
    if l.startswith('{'):
        d=json.loads(l); print(d['value'], d['roofline']['frac'], {k:(v['value'],v['ms_per_step'],v['host_ms_per_step']) for k,v in d['e2e']['variants'].items()})
PY
