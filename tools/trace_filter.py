#!/usr/bin/env python
"""Debug: per-tile globaltimer stamps of filter_compact_kernel on the q2 workload (FLOCKGPU_TRACE=<file>)."""
import os, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
out = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/trace.txt"
os.environ["FLOCKGPU_TRACE"] = out
import numpy as np
import flock_b200 as fb
from flock_b200 import nexgen, plans
ctx = fb.Context(0)
tabs = [ctx.import_batches(nexgen.split_batches(nexgen.bids(10_000_000, seed=s))) for s in (1, 2, 3)]
ec = fb.ExecutionContext(ctx, plans.q2())
for i in range(6):
    ec.feed_tables([tabs[i % 3]])
    n = ec.execute_device(0).num_rows
rows = [l.split() for l in open(out).read().splitlines()]
ntiles, grid = int(rows[0][0]), int(rows[0][1])
a = np.array([[int(x) for x in r] for r in rows[1:1 + ntiles]], dtype=np.int64)
blk_end = np.array([int(r[5]) for r in rows[1:1 + grid]], dtype=np.int64)
t0 = a[:, 0].min()
def stat(x): return f"min {x.min()/1e3:7.2f}  p50 {np.median(x)/1e3:7.2f}  p90 {np.percentile(x,90)/1e3:7.2f}  max {x.max()/1e3:7.2f} us"
print("tiles", ntiles, "grid", grid, "rows kept", n)
print("cta start   ", stat(a[:, 0] - t0))
print("got ticket  ", stat(a[:, 1] - t0))
print("loaded+eval ", stat(a[:, 2] - t0))
print("ranked+lb   ", stat(a[:, 3] - t0))
print("written     ", stat(a[:, 4] - t0))
print("cta end     ", stat(blk_end[blk_end > 0] - t0))
print("load phase  ", stat(a[:, 2] - a[:, 1]))
print("rank+lookbk ", stat(a[:, 3] - a[:, 2]))
print("write phase ", stat(a[:, 4] - a[:, 3]))
sm = a[:, 7]
print("tiles per SM: min", np.bincount(sm).min(), "max", np.bincount(sm).max())
order = np.argsort(a[:, 2])
print("tile index of the 10 last-loaded tiles:", order[-10:], "their load-done times", (a[order[-10:], 2] - t0) / 1e3)
