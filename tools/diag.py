#!/usr/bin/env python
"""tools/diag.py -- small diagnostics behind profiles/: per-repetition q5 times, pageable feed vs staging threads."""
import sys, time, json
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import flock_b200 as fb
from flock_b200 import nexgen, plans

what = sys.argv[1] if len(sys.argv) > 1 else "q5"
ctx = fb.Context(0)
if what == "q5":
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 100_000_000
    bids = ctx.import_batches(nexgen.bids_chunked(n, 42, ["auction"]))
    ec = fb.ExecutionContext(ctx, plans.q5())
    def run():
        ec.feed_tables([bids, bids]); return ec.execute_device(0)
    for _ in range(3): run().num_rows
    for flush in (True, False):
        ts, hs = [], []
        for _ in range(20):
            if flush: ctx.flush_l2()
            ctx.synchronize()
            t = time.perf_counter()
            ctx.timer_start(0); o = run(); ctx.timer_stop(0); o.num_rows
            hs.append((time.perf_counter() - t) * 1e3)
            ts.append(ctx.timer_ms(0))
        print("flush" if flush else "warm", "device ms", [round(x, 3) for x in ts])
        print("      host ms  ", [round(x, 3) for x in hs])
    ctx.profile_begin(); run().num_rows; print(json.dumps(ctx.profile_end()))
    ctx.set_option("host_trace_dump", 1)
    for _ in range(10):
        ctx.flush_l2(); ctx.synchronize(); run().num_rows
    ctx.set_option("host_trace_dump", 1)
elif what == "feed":
    rel = nexgen.bids_chunked(10_000_000, 42)
    ec = fb.ExecutionContext(ctx, plans.q2())
    src = [fb.HostRelation(rel)]
    for thr in (0, 2, 4, 8, 16, 32, 64):
        ctx.set_option("feed_stage_threads", thr)
        ts = []
        for _ in range(8):
            t = time.perf_counter()
            ec.feed_data_sources(src); a = time.perf_counter(); r = ec.execute(); b = time.perf_counter(); ec.clean_data_sources()
            ts.append(((a - t) * 1e3, (b - a) * 1e3, (time.perf_counter() - t) * 1e3))
        print("stage threads", thr, "feed/exec/total ms (median)", [round(float(np.median([x[i] for x in ts[2:]])), 3) for i in range(3)])
ctx.close()
