#!/bin/bash
# round 2, run 25: relaxed window vote v2 (early rebase) A/B; ncu --set full of the q8 and q3 kernels (join_*, gather_*, dense_*, partition_*);
# ncu launch list of the default bench command; one --set full capture of the bench kernel for roofline.traffic
O=gpurun_out/r2_run25; mkdir -p $O
for i in 1 2; do
  timeout 300 python tools/diag.py q5 > $O/diag_q5_strict$i.txt 2>&1
  FLOCKGPU_HIST_RELAXED=1 timeout 300 python tools/diag.py q5 > $O/diag_q5_relaxed$i.txt 2>&1
done
for f in strict1 relaxed1 strict2 relaxed2; do echo "== $f"; grep -E "^warm|agg_hist" $O/diag_q5_$f.txt | cut -c1-200; done
timeout 900 ncu --set full --clock-control none --import-source on --launch-skip 10 --launch-count 10 -f -o $O/q8_kernels python tools/prof_query.py q8 2 > $O/ncu_q8.log 2>&1
tail -3 $O/ncu_q8.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"gather_|join_|filter_" --launch-skip 12 --launch-count 14 -f -o $O/q3_kernels python tools/prof_query.py q3 2 > $O/ncu_q3.log 2>&1
tail -3 $O/ncu_q3.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/bench_launches.csv python bench.py --steps 20 --warmup 3 --no-cpu-baseline --e2e-steps 2 > $O/bench_under_ncu.log 2>&1
tail -2 $O/bench_under_ncu.log | cut -c1-300
timeout 600 ncu --set full --clock-control none --import-source on -k regex:filter_compact -s 4 -c 1 -f -o $O/filter python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-queries --e2e-steps 2 > $O/ncu_filter.log 2>&1
tail -2 $O/ncu_filter.log | cut -c1-300
ls -la $O
