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
elif what == "q8":
    # host-side view of one q8 step over the per-GPU share of the 1 B-event configuration: how many times does the host
    # wait for the device, and for how long (FLOCKGPU_HOST_TRACE=1)
    n_p, n_a, _ = nexgen.relation_counts(125_000_000)
    res = {"person": ctx.import_batches(nexgen.split_batches(nexgen.persons(n_p, 42, 0))),
           "auction": ctx.import_batches(nexgen.split_batches(nexgen.auctions(n_a, 42, 0)))}
    ec = fb.ExecutionContext(ctx, plans.q8())
    def run():
        ec.feed_tables([res[r] for r in plans.SOURCES["q8"]]); return ec.execute_device(0)
    for _ in range(3): run().num_rows
    ts, hs = [], []
    for _ in range(20):
        ctx.synchronize()
        t = time.perf_counter()
        ctx.timer_start(0); o = run(); ctx.timer_stop(0); o.num_rows
        hs.append((time.perf_counter() - t) * 1e3)
        ts.append(ctx.timer_ms(0))
    print("q8 warm device ms", [round(x, 3) for x in ts])
    print("        host ms  ", [round(x, 3) for x in hs])
    ctx.profile_begin(); run().num_rows; print(json.dumps(ctx.profile_end()))
    ctx.set_option("host_trace_dump", 1)
    for _ in range(10):
        ctx.synchronize(); run().num_rows
    ctx.set_option("host_trace_dump", 1)
elif what == "feed":
    # pageable feed of q2's 10 M bids: staging threads x CPU binding.  FOUR distinct host relations in rotation (320 MB:
    # nothing stays in the CPU caches, like bench.py's e2e leg); DIAG_BIND=1 pins the process to the GPU's NUMA node
    # BEFORE the staging threads exist (they inherit the mask)
    import os
    if os.environ.get("DIAG_BIND") == "1":
        import pynvml
        pynvml.nvmlInit(); pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(0))
    print("cpu affinity:", len(os.sched_getaffinity(0)), "cores")
    rels = [nexgen.bids_chunked(10_000_000, 42 + k) for k in range(4)]
    ec = fb.ExecutionContext(ctx, plans.q2())
    srcs = [[fb.HostRelation(r)] for r in rels]
    ctx.set_option("feed_stream_stores", int(os.environ.get("DIAG_STREAM", "1")))
    print("non-temporal staging stores:", os.environ.get("DIAG_STREAM", "1"))
    for thr in [int(x) for x in os.environ.get("DIAG_THREADS", "4,8,12,16,24,32").split(",")]:
        ctx.set_option("feed_stage_threads", thr)
        ts = []
        for i in range(12):
            t = time.perf_counter()
            ec.feed_data_sources(srcs[i % 4]); a = time.perf_counter(); r = ec.execute(); b = time.perf_counter(); ec.clean_data_sources()
            ts.append(((a - t) * 1e3, (b - a) * 1e3, (time.perf_counter() - t) * 1e3))
        print("stage threads", thr, "feed/exec/total ms (median)", [round(float(np.median([x[i] for x in ts[2:]])), 3) for i in range(3)])
    # where does execute() spend its time once the relation is resident?  launch -> survivor count -> export -> release
    parts = []
    for i in range(12):
        ec.feed_data_sources(srcs[i % 4])
        t0 = time.perf_counter(); tab = ec.execute_device(0); t1 = time.perf_counter(); n = tab.num_rows; t2 = time.perf_counter()
        b = tab.to_batch(); t3 = time.perf_counter(); del b, tab; t4 = time.perf_counter()
        ec.clean_data_sources()
        parts.append(((t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3, (t4 - t3) * 1e3))
    print("execute_device / num_rows wait / to_batch (export) / release  ms (median)", [round(float(np.median([x[i] for x in parts[2:]])), 4) for i in range(4)], "rows", n)
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
