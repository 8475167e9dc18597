"""Pins the CPU oracle (`-m "not gpu"`):
  (a) against the reference's own toy goldens, re-expressed as known-answer tests
      (flock/src/runtime/context.rs:428-592, flock/src/launcher/local.rs:169-234,
       flock/src/transmute.rs:319-393);
  (b) against an independent Arrow C++ implementation written from the SQL text (oracle/acero_ref.py);
  (c) against the committed golden vectors (tests/golden/nexmark_golden.json).
NEXMark outputs themselves are unpinned by the reference (every queries/qN.rs only println!s).
"""
import json
from pathlib import Path

import numpy as np
import pyarrow as pa
import pytest

import oracle
from oracle import acero_ref
from flock_b200 import nexgen, plans
from conftest import sources_for

GOLDEN = Path(__file__).parent / "golden" / "nexmark_golden.json"


# ---- (a) the reference's toy goldens -------------------------------------------------------------------
def toy_batch():
    """The 8-row batch of context.rs:441-470 and local.rs:193-220."""
    schema = pa.schema([pa.field("c1", pa.int64(), False), pa.field("c2", pa.float64(), False), pa.field("c3", pa.utf8(), False),
                        pa.field("c4", pa.uint64(), False), pa.field("c5", pa.utf8(), False), pa.field("neg", pa.int64(), False)])
    return pa.RecordBatch.from_arrays([
        pa.array([90, 90, 91, 101, 92, 102, 93, 103], pa.int64()),
        pa.array([92.1, 93.2, 95.3, 96.4, 98.5, 99.6, 100.7, 101.8], pa.float64()),
        pa.array(["a", "a", "d", "b", "b", "d", "c", "c"]),
        pa.array([33, 1, 54, 33, 12, 75, 2, 87], pa.uint64()),
        pa.array(["rapport", "pedantic", "mimesis", "haptic", "baksheesh", "amok", "devious", "c"]),
        pa.array([-90, -90, -91, -101, -92, -102, -93, -103], pa.int64())], schema=schema)


def toy_aggregate_plan(n=8):
    """SELECT MAX(c1), MIN(c2), c3 FROM test WHERE c2 < 99 GROUP BY c3 -- the plan shape of
    flock/src/tests/data/plan/aggregate.json (the reference test adds ORDER BY c3; order is checked below)."""
    b = toy_batch()
    scan = plans.repartition_rr(plans.memory_exec(b.schema, [0, 1, 2]), n)
    pred = plans.binary(plans.column("c2", 1), "Lt", plans.try_cast(plans.literal("Int64", 99), "Float64"))
    filt = plans.coalesce_batches_exec(plans.filter_exec(pred, scan))
    aggrs = [plans.aggregate_expr("max", "MAX(c1)", plans.column("c1", 0), "Int64"),
             plans.aggregate_expr("min", "MIN(c2)", plans.column("c2", 1), "Float64")]
    agg = plans.two_phase_aggregate([("c3", 2)], aggrs, filt, n)
    return plans.projection_exec([(plans.column("MAX(c1)", 1), "MAX(c1)"), (plans.column("MIN(c2)", 2), "MIN(c2)"),
                                  (plans.column("c3", 0), "c3")], agg)


def test_golden_feed_one_data_source():
    # expected table of context.rs:492-500
    out = oracle.execute_plan(toy_aggregate_plan(), [[[toy_batch()]]])
    got = sorted(zip(out["c3"].to_pylist(), out["MAX(c1)"].to_pylist(), out["MIN(c2)"].to_pylist()))
    assert got == [("a", 90, 92.1), ("b", 101, 96.4), ("d", 91, 95.3)]
    assert out.schema.names == ["MAX(c1)", "MIN(c2)", "c3"]


def toy_join_inputs():
    s1 = pa.schema([pa.field("a", pa.utf8(), False), pa.field("b", pa.int32(), False)])
    s2 = pa.schema([pa.field("c", pa.utf8(), False), pa.field("d", pa.int32(), False)])
    b1 = pa.RecordBatch.from_arrays([pa.array(["a", "b", "c", "d"]), pa.array([1, 10, 10, 100], pa.int32())], schema=s1)
    b2 = pa.RecordBatch.from_arrays([pa.array(["a", "b", "c", "d"]), pa.array([1, 10, 10, 100], pa.int32())], schema=s2)
    return b1, b2


def toy_join_plan(n=8):
    """SELECT a, b, d FROM t1 JOIN t2 ON a = c  -- flock/src/tests/data/plan/join.json without its sort/limit."""
    b1, b2 = toy_join_inputs()
    l = plans.coalesce_batches_exec(plans.repartition_hash(plans.repartition_rr(plans.memory_exec(b1.schema, [0, 1]), n), [plans.column("a", 0)], n))
    r = plans.coalesce_batches_exec(plans.repartition_hash(plans.repartition_rr(plans.memory_exec(b2.schema, [0, 1]), n), [plans.column("c", 0)], n))
    j = plans.hash_join_exec(l, r, [(plans.column("a", 0), plans.column("c", 0))])
    return plans.projection_exec([(plans.column("a", 0), "a"), (plans.column("b", 1), "b"), (plans.column("d", 3), "d")],
                                 plans.coalesce_batches_exec(j))


def toy_join_sorted_plan(n=8):
    """The whole of flock/src/tests/data/plan/join.json: ... ORDER BY a LIMIT 3 (global_limit_exec <- sort_exec <- merge_exec)."""
    return plans.global_limit_exec(plans.sort_exec([(plans.column("a", 0), False, False)], plans.coalesce_partitions_exec(toy_join_plan(n))), 3)


def test_golden_join_sort_limit():
    # context.rs:508-592 in full: the expected table of :578-586, order included
    b1, b2 = toy_join_inputs()
    out = oracle.execute_plan(toy_join_sorted_plan(), [[[b1]], [[b2]]])
    assert list(zip(out["a"].to_pylist(), out["b"].to_pylist(), out["d"].to_pylist())) == [("a", 1, 1), ("b", 10, 10), ("c", 10, 10)]


def test_golden_feed_two_data_sources():
    # expected table of context.rs:578-586 (ORDER BY a LIMIT 3 applied here on the host)
    b1, b2 = toy_join_inputs()
    out = oracle.execute_plan(toy_join_plan(), [[[b1]], [[b2]]])
    rows = sorted(zip(out["a"].to_pylist(), out["b"].to_pylist(), out["d"].to_pylist()))
    assert rows[:3] == [("a", 1, 1), ("b", 10, 10), ("c", 10, 10)]
    assert len(rows) == 4


def toy_global_plan(n=8):
    """SELECT MIN(c1), AVG(c4), COUNT(c3) FROM test_table (local.rs:183)."""
    b = toy_batch()
    scan = plans.repartition_rr(plans.memory_exec(b.schema, [0, 2, 3]), n)
    aggrs = [plans.aggregate_expr("min", "MIN(c1)", plans.column("c1", 0), "Int64"),
             plans.aggregate_expr("avg", "AVG(c4)", plans.column("c4", 2), "Float64"),
             plans.aggregate_expr("count", "COUNT(c3)", plans.column("c3", 1), "UInt64")]
    return plans.two_phase_aggregate([], aggrs, scan, n)


def test_golden_local_launcher():
    out = oracle.execute_plan(toy_global_plan(), [[[toy_batch()]]])
    assert out.num_rows == 1
    assert out.column(0).to_pylist() == [90] and out.column(1).to_pylist() == [37.125] and out.column(2).to_pylist() == [8]
    assert [str(f.type) for f in out.schema] == ["int64", "double", "uint64"]


def test_partition_count_invariants():
    # transmute.rs:319-393: RoundRobin(4) of 50 batches -> 13/13/12/12; Hash(8) keeps all 1200 rows
    schema = pa.schema([pa.field("k", pa.int32(), False), pa.field("v", pa.int64(), False)])
    rng = np.random.default_rng(0)
    batches = [pa.RecordBatch.from_arrays([pa.array(rng.integers(0, 100, 24).astype(np.int32)), pa.array(rng.integers(0, 1 << 40, 24))], schema=schema)
               for _ in range(50)]
    ex = oracle.PlanExecutor(plans.repartition_rr(plans.memory_exec(schema, None), 4))
    ex.feed_data_sources([[batches]])
    assert [len(p) for p in ex.execute_partitioned()[0]] == [13, 13, 12, 12]
    ex = oracle.PlanExecutor(plans.repartition_hash(plans.memory_exec(schema, None), [plans.column("k", 0)], 8))
    ex.feed_data_sources([[batches]])
    parts = ex.execute_partitioned()[0]
    assert len(parts) == 8 and sum(b.num_rows for p in parts for b in p) == 1200
    # equal keys meet in one partition
    seen = {}
    for q, p in enumerate(parts):
        for b in p:
            for k in b["k"].to_pylist():
                assert seen.setdefault(k, q) == q


# ---- (b) two independent implementations agree -------------------------------------------------------------
@pytest.mark.parametrize("query", ["q1", "q2", "q3", "q4", "q5", "q6", "q7", "q8"])
@pytest.mark.parametrize("n_parts", [1, 8])
def test_oracle_matches_acero(query, n_parts, events_small):
    got = oracle.execute_plan(plans.QUERIES[query](n_parts), sources_for(query, events_small))
    rels = [events_small[r] for r in dict.fromkeys(plans.SOURCES[query])]
    want = acero_ref.QUERIES[query](*rels)
    oracle.assert_tables_equal(got, want)
    assert got.num_rows > 0


def test_oracle_matches_acero_seed7(events_seed7):
    for query in ["q2", "q3", "q4", "q5", "q6", "q7", "q8"]:
        got = oracle.execute_plan(plans.QUERIES[query](), sources_for(query, events_seed7))
        rels = [events_seed7[r] for r in dict.fromkeys(plans.SOURCES[query])]
        oracle.assert_tables_equal(got, acero_ref.QUERIES[query](*rels))


def test_oracle_threads_agree(events_small):
    a = oracle.execute_plan(plans.q3(), sources_for("q3", events_small), threads=1)
    b = oracle.execute_plan(plans.q3(), sources_for("q3", events_small), threads=4)
    oracle.assert_tables_equal(a, b)


def test_q1_single_rounding():
    # 0.908 * CAST(price AS Float64): exactly one IEEE multiply per row (Appendix C.2)
    bids = nexgen.split_batches(nexgen.bids(65536, seed=1))
    out = oracle.execute_plan(plans.q1(), [[bids]])
    price = pa.Table.from_batches(bids)["price"].to_numpy().astype(np.float64)
    assert np.array_equal(out["price"].to_numpy().view(np.int64), (np.float64(0.908) * price).view(np.int64))


def test_divide_by_zero_is_an_error():
    b = toy_batch()
    pred = plans.binary(plans.binary(plans.column("c1", 0), "Modulo", plans.column("neg", 5)), "Eq", plans.literal("Int64", 0))
    z = pa.RecordBatch.from_arrays([b.column(0), b.column(1), b.column(2), b.column(3), b.column(4), pa.array([0] * 8, pa.int64())], schema=b.schema)
    with pytest.raises(oracle.OracleError, match="Divide by zero"):
        oracle.execute_plan(plans.filter_exec(pred, plans.memory_exec(b.schema, None)), [[[z]]])


def test_empty_and_ragged_inputs():
    bids = nexgen.bids(10_000, seed=3)
    ragged = [bids.slice(0, 1), bids.slice(1, 0), bids.slice(1, 4095), bids.slice(4096, 5904)]
    whole = oracle.execute_plan(plans.q2(), [[[bids]]])
    oracle.assert_tables_equal(oracle.execute_plan(plans.q2(), [[ragged]]), whole)
    empty = oracle.execute_plan(plans.q2(), [[[bids.slice(0, 0)]]])
    assert empty.num_rows == 0 and empty.schema.names == ["auction", "price"]
    # q5 over an empty window: MAX over nothing is NULL, the join keeps nothing
    assert oracle.execute_plan(plans.q5(), [[[bids.slice(0, 0)]], [[bids.slice(0, 0)]]]).num_rows == 0


# ---- (c) committed golden vectors ---------------------------------------------------------------------------
def test_golden_vectors_reproduce():
    g = json.loads(GOLDEN.read_text())
    ev = nexgen.generate(g["n_events"], seed=g["seed"], batch_rows=g["batch_rows"])
    for query, want in g["queries"].items():
        got = oracle.canonical(oracle.execute_plan(plans.QUERIES[query](), sources_for(query, ev)))
        assert got.num_rows == want["num_rows"], query
        assert got.schema.names == want["columns"], query
        head = got.slice(0, len(want["head"])).to_pylist()
        assert json.loads(json.dumps(head, default=str)) == want["head"], query
        assert _digest(got) == want["digest"], query


def _digest(t: pa.Table) -> str:
    import hashlib
    rows = json.dumps(t.to_pylist(), default=str, sort_keys=True)
    return hashlib.sha256(rows.encode()).hexdigest()


def test_native_q8_arm_matches_the_plan_executor_and_an_independent_reference():
    """bench.py's N > 1 CPU arm (oracle.q8_collect -> orc_q8_collect: native threads, Partial DISTINCT -> hash repartition ->
    FinalPartitioned DISTINCT -> partitioned join) against (a) the plan executor running the reference's q8 plan JSON and
    (b) pyarrow's group_by / is_in on inputs with duplicate rows and equal ids under different names."""
    import pyarrow.compute as pc
    ev = nexgen.generate(200_000, seed=21, batch_rows=4096)
    want = oracle.execute_plan(plans.q8(), [[ev[r]] for r in plans.SOURCES["q8"]])
    for parts, threads in ((1, 1), (4, 2), (8, 8), (13, 3)):
        got, times = oracle.q8_collect(ev["person"], ev["auction"], parts, threads, repeat=2)
        assert len(times) == 2
        oracle.assert_tables_equal(got, want)
    rng = np.random.default_rng(3)
    n = 30_000
    pid = rng.integers(0, 12_000, n).astype(np.int32)
    names = np.array(["n%d" % (k % 7) for k in rng.integers(0, 1000, n)])
    persons = [pa.RecordBatch.from_arrays([pa.array(pid[i:i + 4096]), pa.array(names[i:i + 4096])], names=["p_id", "name"]) for i in range(0, n, 4096)]
    sel = rng.integers(0, 18_000, 50_000).astype(np.int32)
    auctions = [pa.RecordBatch.from_arrays([pa.array(sel[i:i + 8192])], names=["seller"]) for i in range(0, 50_000, 8192)]
    distinct = pa.Table.from_batches(persons).group_by(["p_id", "name"], use_threads=False).aggregate([])
    ref = distinct.filter(pc.is_in(distinct["p_id"], value_set=pa.array(np.unique(sel)))).select(["p_id", "name"])
    for parts, threads in ((1, 1), (8, 4)):
        got, _ = oracle.q8_collect(persons, auctions, parts, threads)
        assert got.num_rows == ref.num_rows and got.num_rows < n
        oracle.assert_tables_equal(got, ref)
    assert oracle.q8_collect([persons[0].slice(0, 0)], auctions, 4, 2)[0].num_rows == 0
    assert oracle.q8_collect(persons, [auctions[0].slice(0, 0)], 4, 2)[0].num_rows == 0


def test_native_q5_arm_matches_bincount_and_the_plan_executor():
    """oracle.q5_collect (orc_q5_collect: Partial COUNT per input partition -> hash repartition -> FinalPartitioned COUNT ->
    MAX -> join num = maxn) against numpy.bincount (ties included) and against the plan executor on the reference's q5 plan."""
    bids = nexgen.bids_chunked(1_000_000, 42, ["auction"])
    au = np.concatenate([b["auction"].to_numpy() for b in bids])
    cnt = np.bincount(au)
    winners = np.nonzero(cnt == cnt.max())[0]
    for parts, threads in ((1, 1), (8, 8), (13, 4)):
        got, times = oracle.q5_collect(bids, parts, threads, repeat=2)
        assert len(times) == 2 and sorted(got["auction"].to_pylist()) == winners.tolist() and set(got["num"].to_pylist()) == {int(cnt.max())}
    ties = [pa.RecordBatch.from_arrays([pa.array(np.array([5, 5, 7, 7, 9], np.int32))], names=["auction"])]
    got, _ = oracle.q5_collect(ties, 4, 2)
    assert sorted(zip(got["auction"].to_pylist(), got["num"].to_pylist())) == [(5, 2), (7, 2)]
    assert oracle.q5_collect([ties[0].slice(0, 0)], 4, 2)[0].num_rows == 0
    ev = nexgen.generate(200_000, seed=5, batch_rows=4096)
    want = oracle.execute_plan(plans.q5(), [[ev[r]] for r in plans.SOURCES["q5"]])
    got, _ = oracle.q5_collect(ev["bid"], 8, 4)
    oracle.assert_tables_equal(got, want, check_names=False)


def test_native_q1_and_q3_arms_match_the_plan_executor():
    """oracle.q1_collect / q3_collect (the CPU figures beside bench.py's `queries`) against the plan executor on the
    reference's plans; q3 also on a toy input where equal join keys on both sides give the cross product."""
    ev = nexgen.generate(400_000, seed=42, batch_rows=8192)
    want = oracle.execute_plan(plans.q3(), [[ev[r]] for r in plans.SOURCES["q3"]])
    for parts, threads in ((1, 1), (8, 4), (13, 8)):
        got, times = oracle.q3_collect(ev["auction"], ev["person"], parts, threads, repeat=2)
        assert len(times) == 2
        oracle.assert_tables_equal(got, want)
    a = pa.RecordBatch.from_arrays([pa.array([1, 2, 3, 4], pa.int32()), pa.array([7, 7, 8, 9], pa.int32()), pa.array([10, 10, 11, 10], pa.int32())],
                                   names=["a_id", "seller", "category"])
    p = pa.RecordBatch.from_arrays([pa.array([7, 7, 9, 8], pa.int32()), pa.array(["x", "y", "z", "w"]), pa.array(["c1", "c2", "c3", "c4"]),
                                    pa.array(["or", "id", "wa", "ca"])], names=["p_id", "name", "city", "state"])
    got, _ = oracle.q3_collect([a], [p], 4, 2)
    assert sorted(zip(got["name"].to_pylist(), got["a_id"].to_pylist())) == [("x", 1), ("x", 2), ("y", 1), ("y", 2)]
    assert oracle.q3_collect([a.slice(0, 0)], [p], 4, 2)[0].num_rows == 0
    res, _ = oracle.q1_collect(ev["bid"], 8, 4)
    oracle.assert_tables_equal(pa.Table.from_batches(res), oracle.execute_plan(plans.q1(), [[ev["bid"]]]))
    price = np.concatenate([b["price"].to_numpy() for b in ev["bid"]])
    got = np.concatenate([b["price"].to_numpy() for b in res])
    assert np.array_equal(got.view(np.int64), (np.float64(0.908) * price.astype(np.float64)).view(np.int64))      # one IEEE rounding
