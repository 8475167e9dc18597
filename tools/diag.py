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
elif what == "partition":
    n_p, n_a, _ = nexgen.relation_counts(125_000_000)
    tabs = {"persons(p_id,name)": (ctx.import_batches(nexgen.split_batches(nexgen.persons(n_p, 42, 0, ["p_id", "name"]))), [0]),
            "sellers": (ctx.import_batches(nexgen.split_batches(nexgen.auctions(n_a, 42, 0, ["seller"]))), [0]),
            "bids(4 fixed cols, 10 M)": (ctx.import_batches(nexgen.bids_chunked(10_000_000, 42)), [0])}
    for name, (t, keys) in tabs.items():
        for parts in (2, 8):
            for _ in range(3): ctx.hash_partition(t, keys, parts)
            ctx.profile_begin()
            ts = []
            for _ in range(10):
                ctx.flush_l2(); ctx.timer_start(0); p = ctx.hash_partition(t, keys, parts); ctx.timer_stop(0); ts.append(ctx.timer_ms(0))
            prof = ctx.profile_end()
            print(name, "rows", t.num_rows, "bytes", t.nbytes, "parts", parts, "ms median", round(float(np.median(ts)), 4),
                  {k: round(v["ms"] / v["launches"], 4) for k, v in prof.items()})
ctx.close()
