#!/usr/bin/env python
"""tools/prof_query.py QUERY [REPS] -- runs one NEXMark configuration a few times device-resident (the command ncu wraps)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import flock_b200 as fb
from flock_b200 import nexgen, plans

q = sys.argv[1]
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
ctx = fb.Context(0)
if q == "q5":
    t = ctx.import_batches(nexgen.bids_chunked(100_000_000, 42, ["auction"]))
    tabs = [t, t]
elif q == "q2":
    tabs = [ctx.import_batches(nexgen.bids_chunked(10_000_000, 42, ["auction", "price"]))]
elif q == "q8":
    n_p, n_a, _ = nexgen.relation_counts(125_000_000)
    src = {"person": ctx.import_batches(nexgen.split_batches(nexgen.persons(n_p, 42, 0, ["p_id", "name"]))),
           "auction": ctx.import_batches(nexgen.split_batches(nexgen.auctions(n_a, 42, 0, ["seller"])))}
    tabs = [src[r] for r in plans.SOURCES["q8"]]
elif q == "q3":
    ev = nexgen.generate(10_000_000, seed=42, relations=("person", "auction"), columns={"person": ["p_id", "name", "city", "state"], "auction": ["a_id", "seller", "category"]})
    src = {r: ctx.import_batches(ev[r]) for r in ("auction", "person")}
    tabs = [src[r] for r in plans.SOURCES["q3"]]
ec = fb.ExecutionContext(ctx, plans.QUERIES[q]())
for _ in range(reps):
    ctx.flush_l2()
    ec.feed_tables(tabs)
    print(q, "rows out", ec.execute_device(0).num_rows)
ec.close()
ctx.close()
