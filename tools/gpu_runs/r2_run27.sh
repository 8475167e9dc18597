#!/bin/bash
# round 2, run 27: join slot table at load 0.6 (any capacity); pageable feed: staging threads x NUMA binding
O=gpurun_out/r2_run27; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -x -k "join or q3 or q8 or q4 or q7 or golden" > $O/focus.log 2>&1
tail -3 $O/focus.log
timeout 300 python tools/diag.py q8 > $O/diag_q8.txt 2>&1
grep -E "^q8 warm|join_build" $O/diag_q8.txt | cut -c1-900
timeout 300 python tools/diag.py feed > $O/diag_feed_unbound.txt 2>&1
cat $O/diag_feed_unbound.txt
DIAG_BIND=1 timeout 300 python tools/diag.py feed > $O/diag_feed_bound.txt 2>&1
cat $O/diag_feed_bound.txt
