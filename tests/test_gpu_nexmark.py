"""NEXMark parity on the GPU (`-m gpu`) through the reference-facing surface: the reference's serde-JSON
plan is unmarshalled into an ExecutionContext, fed record batches, executed, cleaned -- exactly the call
sequence of flock-function/src/aws/actor.rs:54-79 -- and the result is compared with the CPU oracle
(canonical sort + bit-exact values, the comparator of flock/src/launcher/aws/mod.rs:675), with the
reference's toy goldens and with the committed golden vectors.  At BASELINE sizes (10 M bids) the checks
are size-independent properties."""
import json
from pathlib import Path

import numpy as np
import pyarrow as pa
import pytest

import flock_b200 as fb
import oracle
from flock_b200 import nexgen, plans
from conftest import sources_for
import test_oracle as goldens

pytestmark = pytest.mark.gpu


def run_gpu(ctx, plan, sources):
    ec = fb.ExecutionContext(ctx, plan)
    ec.feed_data_sources(sources)
    out = ec.execute()
    ec.clean_data_sources()
    ec.close()
    return pa.Table.from_batches(out[0])


@pytest.mark.parametrize("query", ["q1", "q2", "q3", "q4", "q5", "q6", "q7", "q8"])
def test_nexmark_matches_oracle(gpu_ctx, query, events_small):
    got = run_gpu(gpu_ctx, plans.QUERIES[query](), sources_for(query, events_small))
    want = oracle.execute_plan(plans.QUERIES[query](), sources_for(query, events_small))
    assert want.num_rows > 0
    # q1/q2 run in one device partition and filters are stable: even the row order matches the single-partition plan
    oracle.assert_tables_equal(got, want)
    if query in ("q1", "q2"):
        one = oracle.execute_plan(plans.QUERIES[query](1), sources_for(query, events_small))
        oracle.assert_tables_equal(got, one, sort=False)


@pytest.mark.parametrize("query", ["q2", "q3", "q4", "q5", "q6", "q7", "q8"])
def test_nexmark_seed7_full_batches(gpu_ctx, query, events_seed7):
    got = run_gpu(gpu_ctx, plans.QUERIES[query](), sources_for(query, events_seed7))
    oracle.assert_tables_equal(got, oracle.execute_plan(plans.QUERIES[query](), sources_for(query, events_seed7)))


def test_golden_vectors(gpu_ctx):
    g = json.loads(goldens.GOLDEN.read_text())
    ev = nexgen.generate(g["n_events"], seed=g["seed"], batch_rows=g["batch_rows"])
    for query, want in g["queries"].items():
        got = oracle.canonical(run_gpu(gpu_ctx, plans.QUERIES[query](), sources_for(query, ev)))
        assert got.num_rows == want["num_rows"] and got.schema.names == want["columns"], query
        assert goldens._digest(got) == want["digest"], query


def test_reference_toy_goldens(gpu_ctx):
    # context.rs:428-506
    out = run_gpu(gpu_ctx, goldens.toy_aggregate_plan(), [[[goldens.toy_batch()]]])
    assert sorted(zip(out["c3"].to_pylist(), out["MAX(c1)"].to_pylist(), out["MIN(c2)"].to_pylist())) == [("a", 90, 92.1), ("b", 101, 96.4), ("d", 91, 95.3)]
    assert out.schema.names == ["MAX(c1)", "MIN(c2)", "c3"]
    # context.rs:508-592
    b1, b2 = goldens.toy_join_inputs()
    out = run_gpu(gpu_ctx, goldens.toy_join_plan(), [[[b1]], [[b2]]])
    rows = sorted(zip(out["a"].to_pylist(), out["b"].to_pylist(), out["d"].to_pylist()))
    assert rows == [("a", 1, 1), ("b", 10, 10), ("c", 10, 10), ("d", 100, 100)]
    # ... and the fixture in full (global_limit_exec <- sort_exec <- merge_exec on top): ORDER BY a (Utf8) LIMIT 3, order included
    out = run_gpu(gpu_ctx, goldens.toy_join_sorted_plan(), [[[b1]], [[b2]]])
    assert list(zip(out["a"].to_pylist(), out["b"].to_pylist(), out["d"].to_pylist())) == [("a", 1, 1), ("b", 10, 10), ("c", 10, 10)]
    # local.rs:169-234
    out = run_gpu(gpu_ctx, goldens.toy_global_plan(), [[[goldens.toy_batch()]]])
    assert [out.column(i).to_pylist() for i in range(3)] == [[90], [37.125], [8]]


def test_feed_execute_clean_cycle(gpu_ctx, events_small):
    """One context serves many invocations (cloud_context.rs:53-99); unmatched leaves execute empty."""
    ec = fb.ExecutionContext(gpu_ctx, plans.q2())
    want = oracle.execute_plan(plans.q2(1), sources_for("q2", events_small))     # single partition: input order
    for _ in range(3):
        ec.feed_data_sources(sources_for("q2", events_small))
        oracle.assert_tables_equal(pa.Table.from_batches(ec.execute()[0]), want, sort=False)
        ec.clean_data_sources()
        empty = ec.execute()[0][0]
        assert empty.num_rows == 0 and empty.schema.names == ["auction", "price"]
    # person batches do not match a bid leaf (compare_schema, context.rs:402-416): the leaf stays empty
    ec.feed_data_sources([[events_small["person"]]])
    assert ec.execute()[0][0].num_rows == 0
    ec.close()


def test_staged_execution_equals_single_plan(gpu_ctx, events_small):
    """The reference's distributed test shape (launcher/aws/mod.rs:332-471): run q3 stage by stage, transposing
    the shuffle partitions by hand, and compare with the single-plan result."""
    a_plan, p_plan = plans.q3_stage0(4)
    stage0 = fb.ExecutionContext(gpu_ctx, [a_plan, p_plan])
    assert stage0.is_shuffling()
    stage0.feed_data_sources(sources_for("q3", events_small))
    a_parts, p_parts = stage0.execute_partitioned()
    assert len(a_parts) == 4 and len(p_parts) == 4
    b = events_small
    join = plans.projection_exec(
        [(plans.column("name", 4), "name"), (plans.column("city", 5), "city"), (plans.column("state", 6), "state"), (plans.column("a_id", 0), "a_id")],
        plans.coalesce_batches_exec(plans.hash_join_exec(
            plans.memory_exec(a_parts[0][0].schema, None), plans.memory_exec(p_parts[0][0].schema, None),
            [(plans.column("seller", 1), plans.column("p_id", 0))])))
    stage1 = fb.ExecutionContext(gpu_ctx, join)
    outs = []
    for q in range(4):                                   # partition q of both sides meets in function q
        stage1.feed_data_sources([[a_parts[q]], [p_parts[q]]])
        outs += stage1.execute()[0]
        stage1.clean_data_sources()
    staged = pa.Table.from_batches(outs)
    single = run_gpu(gpu_ctx, plans.q3(), sources_for("q3", b))
    oracle.assert_tables_equal(staged, single)
    oracle.assert_tables_equal(single, oracle.execute_plan(plans.q3(), sources_for("q3", b)))


def test_zero_copy_feed_of_pinned_batches(gpu_ctx):
    """feed_zero_copy: page-locked, uniformly batched columns are read in place by the vectorised filter (q2) and
    copied to HBM on first use by everything else (q5, q1); pageable or ragged batches silently take the copy path."""
    bids = nexgen.split_batches(nexgen.bids(5 * 65536 + 1234, seed=17), 65536)
    pinned = [gpu_ctx.pinned_copy(b) for b in bids]
    gpu_ctx.set_option("feed_zero_copy", 1)
    try:
        for query in ("q2", "q5", "q1"):
            for batches in (pinned, bids, [bids[0].slice(0, 1000)] + pinned[1:]):
                src = [[batches]] * len(plans.SOURCES[query])
                got = run_gpu(gpu_ctx, plans.QUERIES[query](), src)
                want = oracle.execute_plan(plans.QUERIES[query](1), src)
                oracle.assert_tables_equal(got, want, sort=query != "q2")
        # a filter with a computed output column cannot run in place: it must densify and still be right
        t = gpu_ctx.import_batches(pinned)
        ec = fb.ExecutionContext(gpu_ctx, plans.projection_exec(
            [(plans.column("auction", 0), "auction"), (plans.binary(plans.column("price", 1), "Plus", plans.literal("Int32", 1)), "p1")],
            plans.filter_exec(plans.binary(plans.cast(plans.column("auction", 0), "Int64"), "Gt", plans.literal("Int64", 1500)),
                              plans.memory_exec(nexgen.bid_schema(), [0, 2]))))
        ec.feed_data_sources([fb.HostRelation(pinned)])
        got = pa.Table.from_batches(ec.execute()[0])
        ec.close()
        full = pa.Table.from_batches(bids)
        keep = full["auction"].to_numpy() > 1500
        assert np.array_equal(got["auction"].to_numpy(), full["auction"].to_numpy()[keep])
        assert np.array_equal(got["p1"].to_numpy(), full["price"].to_numpy()[keep] + 1)
    finally:
        gpu_ctx.set_option("feed_zero_copy", 0)


def test_hopping_windows_assembled_on_the_device(gpu_ctx):
    """q5 over Hopping(size 4, hop 2) windows (benchmarks/src/nexmark/main.rs:116-123 runs it as Hopping(10 s, 5 s)): every
    epoch is uploaded ONCE, windows are concatenated from resident epochs (hopping.rs:54-74), and each window's hot items
    equal the oracle's over the same epochs.  hop = size is the tumbling case (q7 / q8)."""
    epochs = [nexgen.bids(20_000 + 1000 * e, seed=70 + e, first_bid=100_000 * e) for e in range(9)]
    ec = fb.ExecutionContext(gpu_ctx, plans.q5())
    for size, hop in ((4, 2), (3, 3)):
        w = fb.Window(gpu_ctx, size, hop)
        seen = []
        for e, b in enumerate(epochs):
            w.push(gpu_ctx.import_batches([b]))
            while w.ready:
                t, first = w.next()
                seen.append(first)
                span = epochs[first:first + size]
                assert t.num_rows == sum(x.num_rows for x in span)
                assert t.to_batch().equals(pa.Table.from_batches(span).combine_chunks().to_batches()[0])
                ec.feed_tables([t, t])
                got = ec.execute_device(0).to_arrow()
                oracle.assert_tables_equal(got, oracle.execute_plan(plans.q5(), [[span], [span]]))
        assert seen == list(range(0, len(epochs) - size + 1, hop))
        w.close()
    ec.close()
    with pytest.raises(fb.FlockGpuError):
        fb.Window(gpu_ctx, 2, 3)                  # hop > size: rejected like hopping.rs:38-43


def test_device_resident_feed(gpu_ctx, events_small):
    bids = gpu_ctx.import_batches(events_small["bid"])
    ec = fb.ExecutionContext(gpu_ctx, plans.q5())
    ec.feed_tables([bids, bids])
    got = ec.execute_device(0).to_arrow()
    oracle.assert_tables_equal(got, oracle.execute_plan(plans.q5(), sources_for("q5", events_small)))
    ec.close()


# ---- BASELINE sizes: size-independent properties ------------------------------------------------------------------
@pytest.fixture(scope="module")
def bids_10m():
    return nexgen.split_batches(nexgen.bids(10_000_000, seed=42))        # 152 full 64 Ki-row batches + 1 short


def test_q2_full_size_properties(gpu_ctx, bids_10m):
    t = gpu_ctx.import_batches(bids_10m, projection=[0, 2])
    ec = fb.ExecutionContext(gpu_ctx, plans.q2())
    ec.feed_tables([gpu_ctx.import_batches(bids_10m)])
    out = ec.execute_device(0)
    got = out.to_arrow()
    auction = np.concatenate([b["auction"].to_numpy() for b in bids_10m])
    price = np.concatenate([b["price"].to_numpy() for b in bids_10m])
    keep = np.fmod(auction.astype(np.int64), 123) == 0
    assert got.num_rows == int(keep.sum())                                  # count
    assert np.array_equal(got["auction"].to_numpy(), auction[keep])         # stable order, exact values
    assert np.array_equal(got["price"].to_numpy(), price[keep])
    # idempotence: filtering the output again keeps everything
    again = gpu_ctx.filter_project(out, fb.col(0).cast("int64") % 123 == 0)
    assert again.num_rows == got.num_rows
    # complement: the negated predicate keeps exactly the other rows
    rest = gpu_ctx.filter_project(t, ~(fb.col(0).cast("int64") % 123 == 0))
    assert rest.num_rows + got.num_rows == auction.size
    ec.close()


def test_q5_full_size_against_bincount(gpu_ctx, bids_10m):
    ec = fb.ExecutionContext(gpu_ctx, plans.q5())
    bids = gpu_ctx.import_batches(bids_10m, projection=[0])
    ec.feed_tables([bids, bids])
    got = ec.execute_device(0).to_arrow()
    auction = np.concatenate([b["auction"].to_numpy() for b in bids_10m])
    counts = np.bincount(auction)
    winners = np.nonzero(counts == counts.max())[0]
    assert sorted(got["auction"].to_pylist()) == winners.tolist() and set(got["num"].to_pylist()) == {int(counts.max())}
    # checksum of checksums: the per-auction counts sum to the number of bids
    per = gpu_ctx.hash_aggregate(bids, [0], [("count", -1, "n")], "single").to_arrow()
    assert per.num_rows == int((counts > 0).sum()) and sum(per["n"].to_pylist()) == auction.size
    ec.close()
