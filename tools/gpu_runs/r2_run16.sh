#!/bin/bash
mkdir -p gpurun_out/r2_run16
for t in 1 0; do
  echo "== FLOCKGPU_HIST_TMA=$t (2 stages, 3 CTAs/SM)"
  FLOCKGPU_HIST_TMA=$t timeout 200 python tools/diag.py q5 > gpurun_out/r2_run16/diag_q5_tma$t.txt 2>&1
  head -1 gpurun_out/r2_run16/diag_q5_tma$t.txt | cut -c1-170; grep agg_ gpurun_out/r2_run16/diag_q5_tma$t.txt
done
