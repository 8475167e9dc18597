#!/bin/bash
# round 2, run 37: what execute() costs after a pageable feed (launch, wait, export, release), with host spans
O=gpurun_out/r2_run37; mkdir -p $O
FLOCKGPU_HOST_TRACE=1 DIAG_THREADS=8 timeout 300 python tools/diag.py feed > $O/feed_parts.txt 2>&1; cat $O/feed_parts.txt | tail -25
