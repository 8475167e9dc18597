#!/bin/bash
# round 2, run 40: the whole -m gpu suite and smoke() on the last commit (non-temporal staging in every pageable import)
O=gpurun_out/r2_run40; mkdir -p $O
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/smoke.txt 2>&1; tail -1 $O/smoke.txt
timeout 900 python -m pytest tests -m gpu -q > $O/pytest.log 2>&1; tail -3 $O/pytest.log
