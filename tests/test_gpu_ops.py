"""Operator-level parity on the GPU (`-m gpu`), every call through the C ABI (ctypes), checked against the
CPU oracle on the same seeded inputs.  Bit-exact: integers, bytes and indices everywhere; the only
floating-point kernels (q1's 0.908 * CAST(price AS Float64) and AVG) are also compared bit for bit
because each value is produced by a single IEEE operation."""
import numpy as np
import pyarrow as pa
import pyarrow.compute as pc
import pytest

import flock_b200 as fb
import oracle
from flock_b200 import col, lit, nexgen, plans, sharding

pytestmark = pytest.mark.gpu


def rb(**cols):
    return pa.RecordBatch.from_arrays(list(cols.values()), names=list(cols.keys()))


def mixed_batch(n, seed):
    rng = np.random.default_rng(seed)
    words = ["", "a", "or", "id", "ca", "portland", "san francisco", "x" * 37, "émile", "日本"]
    return pa.RecordBatch.from_arrays([
        pa.array(rng.integers(-1000, 1000, n).astype(np.int32)),
        pa.array(rng.integers(-(1 << 50), 1 << 50, n)),
        pa.array(rng.integers(0, 1 << 62, n).astype(np.uint64)),
        pa.array(rng.normal(0, 1e6, n)),
        pa.array(rng.integers(1_436_918_400_000, 1_436_918_500_000, n), pa.timestamp("ms")),
        pa.array([words[k] for k in rng.integers(0, len(words), n)]),
        pa.array(rng.integers(0, 50, n).astype(np.int32)),
    ], schema=pa.schema([pa.field("i32", pa.int32(), False), pa.field("i64", pa.int64(), False), pa.field("u64", pa.uint64(), False),
                         pa.field("f64", pa.float64(), False), pa.field("ts", pa.timestamp("ms"), False), pa.field("s", pa.utf8(), False),
                         pa.field("k", pa.int32(), False)], metadata={"name": "mixed"}))


# ---- host <-> HBM ---------------------------------------------------------------------------------------
@pytest.mark.parametrize("n", [0, 1, 31, 4097, 100_003])
def test_import_export_roundtrip(gpu_ctx, n):
    b = mixed_batch(n, seed=n)
    pieces = [b] if n < 10 else [b.slice(0, 7), b.slice(7, 0), b.slice(7, n // 2), b.slice(7 + n // 2)]   # ragged + sliced (offset != 0)
    t = gpu_ctx.import_batches(pieces)
    assert t.num_rows == n and t.num_columns == 7
    out = t.to_batch()
    assert out.schema.equals(b.schema, check_metadata=True)
    assert out.equals(b)
    if n > 100:
        assert t.to_batch(50, 40).equals(b.slice(50, 40))          # Utf8 offsets are rebased on export
        proj = gpu_ctx.import_batches(pieces, projection=[5, 0])
        assert proj.to_batch().equals(b.select(["s", "i32"]))
        both = gpu_ctx.concat([t, proj]) if False else gpu_ctx.concat([t, t])
        assert both.to_batch().equals(pa.Table.from_batches([b, b]).combine_chunks().to_batches()[0])


def _ipc_frame(batch: pa.RecordBatch):
    """(data_header, data_body) of the batch as arrow-rs / pyarrow write it (the DataFrame of flock's Payload)."""
    msg = pa.ipc.read_message(batch.serialize())
    return msg.metadata, msg.body


def _read_frame(schema: pa.Schema, header: bytes, body: bytes) -> pa.RecordBatch:
    import struct
    stream = struct.pack("<Ii", 0xFFFFFFFF, len(header)) + header + body       # encapsulated message (header is 8-byte padded)
    return pa.ipc.read_record_batch(pa.ipc.read_message(stream), schema)


@pytest.mark.parametrize("n", [0, 1, 1000, 70_001])
def test_ipc_frames_roundtrip(gpu_ctx, n):
    """Payload frames in (pyarrow writes them) = the table a batch import gives; frames out are read back by pyarrow's own
    IPC reader.  Byte-for-byte the buffers are the column buffers."""
    b = mixed_batch(n, seed=n + 3)
    pieces = [b] if n < 10 else [b.slice(0, n // 3), b.slice(n // 3)]
    t = gpu_ctx.import_ipc(b.schema, [_ipc_frame(p) for p in pieces])
    assert t.num_rows == n and t.to_batch().equals(b)
    proj = gpu_ctx.import_ipc(b.schema, [_ipc_frame(p) for p in pieces], projection=[5, 0])
    assert proj.to_batch().equals(b.select(["s", "i32"]))
    header, body = t.to_ipc()
    assert _read_frame(b.schema, header, body).equals(b)
    assert len(header) % 8 == 0 and len(body) % 8 == 0
    if n > 100:
        header, body = t.to_ipc(37, 50)                                          # Utf8 offsets rebased, like arrow-rs
        assert _read_frame(b.schema, header, body).equals(b.slice(37, 50))
        # a frame our writer produced is a frame our reader takes
        again = gpu_ctx.import_ipc(b.schema, [t.to_ipc(0, 64), t.to_ipc(64, -1)])
        assert again.to_batch().equals(b)
    with pytest.raises(fb.FlockGpuError):
        gpu_ctx.import_ipc(b.schema, [(b"\x00" * 4, b"")])


def _ndjson(batch: pa.RecordBatch, shuffle_keys: bool = False, seed: int = 0) -> bytes:
    """The lines serde_json writes for NEXMark events: one compact object per row (timestamps as integer milliseconds)."""
    import json, random
    cols = {}
    for f in batch.schema:
        c = batch.column(f.name)
        cols[f.name] = c.cast(pa.int64()).to_pylist() if pa.types.is_timestamp(f.type) else c.to_pylist()
    rnd = random.Random(seed)
    lines = []
    for i in range(batch.num_rows):
        keys = list(cols)
        if shuffle_keys:
            rnd.shuffle(keys)
        lines.append(json.dumps({k: cols[k][i] for k in keys}, separators=(",", ":"), ensure_ascii=bool(i & 1)))
    return ("\n".join(lines) + ("\n" if lines else "")).encode("utf-8")


@pytest.mark.parametrize("relation", ["bid", "person", "auction"])
def test_ndjson_events_to_table(gpu_ctx, relation):
    """event_bytes_to_batch (transmute.rs:255-266): NEXMark events as NDJSON parse into exactly the batch they came from."""
    ev = nexgen.generate(60_000, seed=5, batch_rows=1 << 20)
    b = ev[relation][0]
    assert b.num_rows > 1000
    got = gpu_ctx.import_ndjson(b.schema, _ndjson(b)).to_batch()
    assert got.equals(b)
    got = gpu_ctx.import_ndjson(b.schema, _ndjson(b, shuffle_keys=True, seed=3)).to_batch()      # fields are found by name
    assert got.equals(b)
    part = pa.schema([b.schema.field(i) for i in (b.num_columns - 1, 0)], metadata=b.schema.metadata)  # the rest is skipped
    assert gpu_ctx.import_ndjson(part, _ndjson(b)).to_batch().equals(b.select([b.num_columns - 1, 0]))


def test_ndjson_syntax_corners_match_arrow(gpu_ctx):
    import pyarrow.json as pj, io
    schema = pa.schema([("i", pa.int32()), ("l", pa.int64()), ("u", pa.uint64()), ("f", pa.float64()), ("s", pa.utf8()), ("t", pa.timestamp("ms"))])
    arrow_schema = pa.schema([f if f.name != "t" else pa.field("t", pa.int64()) for f in schema])      # Arrow C++ wants timestamp STRINGS; epochs go in as integers
    lines = [
        r'{"i":-2147483648,"l":-9223372036854775808,"u":18446744073709551615,"f":-0.5,"s":"","t":0}',
        r'{ "s" : "a\"b\\c\/d\n\té日😀" , "i":7, "extra":{"x":[1,{"y":"}"}],"z":null}, "l":1,"u":0,"f":1e3,"t":1436918400000}',
        '{"i":2147483647,"l":9223372036854775807,"u":1,"f":123456789012345,"s":"é日 plain utf-8","t":-1,"tail":[true,false]}',
        r'{"i":1,"l":1,"u":1,"f":1,"s":"\u00e9\u65e5\ud83d\ude00 \u0041","t":1}',
        '{"f":2.5E-3,"s":"x","i":0,"l":0,"u":0,"t":5}\r',
    ]
    text = ("\n".join(lines) + "\n").encode("utf-8")
    want = pj.read_json(io.BytesIO(text), parse_options=pj.ParseOptions(explicit_schema=arrow_schema, unexpected_field_behavior="ignore"))
    got = gpu_ctx.import_ndjson(schema, text).to_arrow()
    assert got.equals(want.select(got.schema.names).cast(got.schema))
    assert got["s"].to_pylist()[1] == 'a"b\\c/d\n\té日😀'
    assert gpu_ctx.import_ndjson(schema, b"").num_rows == 0
    assert gpu_ctx.import_ndjson(schema, text[:-1]).num_rows == 5                      # no newline behind the last line
    for bad, why in ((b'{"i":1}\n', "missing"), (b'{"i":1.5,"l":1,"u":1,"f":1,"s":"","t":1}\n', "type"), (b'{"i":3000000000,"l":1,"u":1,"f":1,"s":"","t":1}\n', "range"),
                     (b'{"i":1,"l":1,"u":1,"f":1,"s":null,"t":1}\n', "null"), (b'{"i":1,"l":1,"u":1,"f":1,"s":"abc,"t":1}\n', "malformed"),
                     (b'{"i":1,"l":1,"u":1,"f":0.1234567890123456789,"s":"","t":1}\n', "significant")):
        with pytest.raises(fb.FlockGpuError) as info:
            gpu_ctx.import_ndjson(schema, text + bad)
        assert info.value.code == -5 and "line 6" in info.value.message and why in info.value.message, info.value.message


def test_import_rejects_unknown_types(gpu_ctx):
    with pytest.raises(fb.FlockGpuError) as info:
        gpu_ctx.import_batches([rb(x=pa.array([1.0, 2.0], pa.float32()))])
    assert info.value.code == -2


# ---- NULLs: validity travels with the rows, and every operator follows SQL / DataFusion semantics (App. C.3 / C.7 / C.8) --------
def nullable_batch(n, seed, null_every=(5, 7, 3, 11)):
    rng = np.random.default_rng(seed)
    words = ["", "a", "or", "id", "portland", "x" * 21, "日本"]
    def holes(arr, k, offset=0):
        mask = (np.arange(n) + offset) % k == 0
        return pa.array(arr, mask=mask) if isinstance(arr, np.ndarray) else pa.array([None if m else v for v, m in zip(arr, mask)])
    return pa.RecordBatch.from_arrays([
        holes(rng.integers(0, 12, n).astype(np.int32), null_every[0]),                      # g: group / join key with NULLs
        holes(rng.integers(-1000, 1000, n), null_every[1], 2),                             # v: Int64 with NULLs
        holes(rng.normal(0, 50, n).round(0), null_every[2], 1),                            # f: Float64 (integral values: exact sums)
        holes([words[k] for k in rng.integers(0, len(words), n)], null_every[3], 4),       # s: Utf8 with NULLs
        pa.array(np.arange(n, dtype=np.int32)),                                            # id: no NULLs
    ], names=["g", "v", "f", "s", "id"])


@pytest.mark.parametrize("n", [1, 9, 5000, 150_001])
def test_nulls_roundtrip_filter_take(gpu_ctx, n):
    b = nullable_batch(n, n)
    pieces = [b] if n < 20 else [b.slice(0, 3), b.slice(3, n // 2), b.slice(3 + n // 2)]          # bit offsets that are not byte aligned
    t = gpu_ctx.import_batches(pieces)
    assert t.to_batch().equals(b) and t.to_batch(min(1, n - 1), max(n - 3, 0)).equals(b.slice(min(1, n - 1), max(n - 3, 0)))
    # frames with validity in, frames with validity out
    h, body = t.to_ipc()
    assert _read_frame(b.schema, h, body).equals(b)
    assert gpu_ctx.import_ipc(b.schema, [_ipc_frame(p) for p in pieces]).to_batch().equals(b)
    # FilterExec: a NULL predicate drops the row (three-valued AND / OR / NOT), NULL payload survives as NULL
    cases = [
        (col(1) > 0, pc.greater(b["v"], 0)),
        ((col(1) > 0) | (col(0) == 3), pc.or_kleene(pc.greater(b["v"], 0), pc.equal(b["g"], 3))),
        ((col(1) > 0) & ~(col(2) < 10.0), pc.and_kleene(pc.greater(b["v"], 0), pc.invert(pc.less(b["f"], 10.0)))),
        ((col(3) == "or") | (col(4) % 2 == 0), pc.or_kleene(pc.equal(b["s"], "or"), pc.equal(pc.bit_wise_and(b["id"], 1), 0))),
        (~(col(0) == 3), pc.invert(pc.equal(b["g"], 3))),
    ]
    for pred, mask in cases:
        got = gpu_ctx.filter_project(t, pred).to_batch()
        want = b.filter(mask.fill_null(False))
        assert got.equals(want)
    # take through a selection, then concat: validity is one more column
    kept = gpu_ctx.filter_project(t, col(4) % 3 == 0)
    both = gpu_ctx.concat([kept, t]).to_batch()
    assert both.equals(pa.Table.from_batches([b.filter(pc.equal(pc.subtract(b["id"], pc.multiply(pc.divide(b["id"], 3), 3)), 0)), b]).combine_chunks().to_batches()[0])
    # ProjectionExec: a computed value is NULL where an input is NULL
    proj = gpu_ctx.filter_project(t, None, [col(1) * 2 + col(4), col(2) * 0.5, col(4)], ["a", "h", "id"]).to_batch()
    assert proj["a"].equals(pc.add(pc.multiply(b["v"], 2), pc.cast(b["id"], pa.int64()))) and proj["h"].equals(pc.multiply(b["f"], 0.5))


@pytest.mark.parametrize("n", [7, 4000, 300_017])
def test_nulls_aggregate_join_partition(gpu_ctx, n):
    b = nullable_batch(n, n + 1)
    t = gpu_ctx.import_batches([b])
    tbl = pa.Table.from_batches([b])
    # HashAggregateExec: NULL arguments are skipped, a group of only NULLs yields NULL, the NULL key is a group
    key = lambda tb, names: sorted(zip(*[tb[c].to_pylist() for c in names]), key=lambda r: (r[0] is None, r[0] if r[0] is not None else 0))
    for aggs, ref, ref_names in (
            ([("count", -1, "n"), ("count", 1, "nv"), ("sum", 1, "sv"), ("max", 1, "xv")],
             [([], "count_all"), ("v", "count"), ("v", "sum"), ("v", "max")], ["count_all", "v_count", "v_sum", "v_max"]),
            ([("min", 2, "mf"), ("avg", 2, "af"), ("count", 2, "nf")],
             [("f", "min"), ("f", "mean"), ("f", "count")], ["f_min", "f_mean", "f_count"])):
        want = tbl.group_by("g", use_threads=False).aggregate(ref)
        for mode_path in ("single", "two_phase"):
            if mode_path == "single":
                got = gpu_ctx.hash_aggregate(t, [0], aggs, "single").to_arrow()
            else:
                part = gpu_ctx.hash_aggregate(t, [0], aggs, "partial")
                state, fin = 1, []
                for f, _, name in aggs:
                    fin.append((f, state, name))
                    state += 2 if f == "avg" else 1
                got = gpu_ctx.hash_aggregate(part, [0], fin, "final_partitioned").to_arrow()
            assert key(got, ["g"] + [a[2] for a in aggs]) == key(want, ["g"] + ref_names), (aggs, mode_path)
    # global aggregate over a column that is NULL throughout
    allnull = gpu_ctx.import_batches([rb(x=pa.array([None] * 5, pa.int64()), y=pa.array([1, 2, 3, 4, 5]))])
    one = gpu_ctx.hash_aggregate(allnull, [], [("max", 0, "m"), ("count", 0, "c"), ("count", -1, "n"), ("sum", 1, "s")], "single").to_arrow()
    assert [one.column(i).to_pylist() for i in range(4)] == [[None], [0], [5], [15]]
    # HashJoinExec: NULL keys match nothing (not even each other); NULL payload columns come through
    r = nullable_batch(min(max(n // 3, 4), 40), n + 2, null_every=(4, 5, 6, 7))
    r = r.rename_columns(["g2", "v2", "f2", "s2", "id2"])
    got = gpu_ctx.hash_join(t, gpu_ctx.import_batches([r]), [0], [0]).to_arrow()
    want = tbl.join(pa.Table.from_batches([r]), keys="g", right_keys="g2", join_type="inner", coalesce_keys=False, use_threads=False)
    oracle.assert_tables_equal(got, want.select(got.schema.names), check_names=False)
    assert got["g"].null_count == 0 and (n < 100 or got["v"].null_count > 0)
    got2 = gpu_ctx.hash_join(t, gpu_ctx.import_batches([r]), [3, 0], [3, 0]).to_arrow()                      # Utf8 + Int32 key, both with NULLs
    want2 = tbl.join(pa.Table.from_batches([r]), keys=["s", "g"], right_keys=["s2", "g2"], join_type="inner", coalesce_keys=False, use_threads=False)
    oracle.assert_tables_equal(got2, want2.select(got2.schema.names), check_names=False)
    # RepartitionExec(Hash): rows with a NULL key land together; validity moves with the rows; nothing is lost
    parts = [p.to_batch() for p in gpu_ctx.hash_partition(t, [0], 5)]
    assert sum(p.num_rows for p in parts) == n
    home = {}
    for q, p in enumerate(parts):
        for g in set(p["g"].to_pylist()):
            assert home.setdefault(g, q) == q
    oracle.assert_tables_equal(pa.Table.from_batches(parts), tbl)
    # operators without NULL semantics refuse loudly
    with pytest.raises(fb.FlockGpuError) as info:
        gpu_ctx.sort(t, [1])
    assert info.value.code == -2 and "NULL" in info.value.message


# ---- FilterExec / ProjectionExec ----------------------------------------------------------------------------
PREDICATES = [
    lambda: col(0).cast("int64") % 123 == 0,
    lambda: col(0).cast("int64") % 7 == -3,
    lambda: col(0).cast("int64") == 10,
    lambda: col(0).cast("int64") < 0,
    lambda: (col(5) == "or") | (col(5) == "id") | (col(5) == "ca"),
    lambda: col(5) >= "id",
    lambda: (col(1) > 0) & ~(col(3) < lit(0).cast("float64")),
    lambda: (col(4) >= lit(1_436_918_450_000, "timestamp")) & (col(2) > lit(1 << 61, "uint64")),
    lambda: (col(1) + col(0).cast("int64") * 3) % 1000 == 1,
    lambda: ~(col(0).cast("int64") % 123 == 0),          # negated single term: the vectorised kernel with the inverted comparison
    lambda: col(0).cast("int64") > 5000,        # nothing survives
    lambda: col(0).cast("int64") >= -5000,      # everything survives
]


def oracle_filter(b, pred_e, proj_es=None, names=None):
    """Reference result through the oracle's own evaluator, driven by the same postfix tokens."""
    import ctypes as C
    cols = oracle.Cols(b)
    toks, keep = oracle._ctoks(pred_e.tokens)
    mask = np.zeros(max(b.num_rows, 1), np.uint8)
    oracle._check(oracle.lib().orc_eval_predicate(cols.arr, b.num_columns, C.c_int64(b.num_rows), toks, len(toks), mask.ctypes.data_as(C.c_void_p)), "pred")
    return b.filter(pa.array(mask[:b.num_rows].astype(bool)))


@pytest.mark.parametrize("case", range(len(PREDICATES)))
@pytest.mark.parametrize("n", [1, 4096, 70_001])
def test_filter_matches_oracle(gpu_ctx, case, n):
    b = mixed_batch(n, seed=100 + case)
    pred = PREDICATES[case]()
    t = gpu_ctx.import_batches([b.slice(0, n // 3), b.slice(n // 3)])
    got = gpu_ctx.filter_project(t, pred).to_batch()
    want = oracle_filter(b, pred)
    assert got.num_rows == want.num_rows
    assert got.equals(want)                                   # stable: surviving rows keep their input order


@pytest.mark.parametrize("m", [1, 2, 3, 96, 123, 1000, 1 << 20, 3 << 29, (1 << 31) - 1, -123])
def test_divisibility_fast_path_all_moduli(gpu_ctx, m):
    """`CAST(i32 AS Int64) % m = 0` runs the one-multiply divisibility test (PredI32 MODE 2): odd, even and
    power-of-two moduli, negative dividends and INT_MIN; `!= 0` and `% m = c` take the Lemire path (MODE 1)."""
    rng = np.random.default_rng(abs(m))
    n = 200_003
    x = rng.integers(-(1 << 31), 1 << 31, n).astype(np.int32)
    am = abs(m)
    x[::7] = (rng.integers(-((1 << 31) // am), ((1 << 31) - 1) // am + 1, len(x[::7])) * am).astype(np.int32)   # plenty of multiples
    x[:4] = [0, -(1 << 31), (1 << 31) - 1, -1]
    b = rb(v=pa.array(x), i=pa.array(np.arange(n, dtype=np.int64)))
    t = gpu_ctx.import_batches([b])
    rem = np.fmod(x.astype(np.int64), np.int64(m))            # sign follows the dividend, like Rust / Arrow
    for pred, keep in ((col(0).cast("int64") % m == 0, rem == 0), (col(0).cast("int64") % m != 0, rem != 0),
                       (col(0).cast("int64") % m == 1, rem == 1)):
        got = gpu_ctx.filter_project(t, pred).to_batch()
        assert np.array_equal(got["i"].to_numpy(), np.nonzero(keep)[0])
        assert got.equals(oracle_filter(b, pred))


def test_filter_with_projection_and_computed_columns(gpu_ctx):
    b = mixed_batch(50_000, seed=9)
    t = gpu_ctx.import_batches([b])
    got = gpu_ctx.filter_project(t, col(0).cast("int64") % 5 == 0,
                                 [col(5), col(0), 0.908 * col(0).cast("float64"), col(1) - col(0).cast("int64"), col(4)],
                                 ["s", "i32", "scaled", "diff", "ts"]).to_batch()
    keep = np.fmod(b["i32"].to_numpy().astype(np.int64), 5) == 0
    w = b.filter(pa.array(keep))
    i32 = w["i32"].to_numpy()
    assert got.schema.names == ["s", "i32", "scaled", "diff", "ts"]
    assert got["s"].equals(w["s"]) and got["i32"].equals(w["i32"]) and got["ts"].equals(w["ts"])
    assert np.array_equal(got["scaled"].to_numpy().view(np.int64), (np.float64(0.908) * i32.astype(np.float64)).view(np.int64))
    assert np.array_equal(got["diff"].to_numpy(), w["i64"].to_numpy() - i32)


def test_projection_only_is_zero_copy_and_exact(gpu_ctx):
    bids = nexgen.bids(65536 + 3, seed=11)
    t = gpu_ctx.import_batches([bids])
    k0 = gpu_ctx.kernel_launches
    out = gpu_ctx.filter_project(t, None, [col(0), col(1), 0.908 * col(2).cast("float64"), col(3)], ["auction", "bidder", "price", "b_date_time"])
    assert gpu_ctx.kernel_launches - k0 == 1                  # only the computed column costs a kernel
    got = out.to_batch()
    price = bids["price"].to_numpy().astype(np.float64)
    assert np.array_equal(got["price"].to_numpy().view(np.int64), (np.float64(0.908) * price).view(np.int64))
    assert got["auction"].equals(bids["auction"]) and got["b_date_time"].equals(bids["b_date_time"])
    gen = gpu_ctx.filter_project(t, None, [(col(2) + 7) * 3, col(2).cast("int64") % 1000], ["a", "b"]).to_batch()   # generic interpreter
    p = bids["price"].to_numpy()
    assert np.array_equal(gen["a"].to_numpy(), ((p + 7) * 3).astype(np.int32)) and np.array_equal(gen["b"].to_numpy(), p.astype(np.int64) % 1000)


def test_divide_by_zero_is_reported(gpu_ctx):
    b = rb(a=pa.array([1, 2, 3], pa.int64()), z=pa.array([1, 0, 2], pa.int64()))
    t = gpu_ctx.import_batches([b])
    with pytest.raises(fb.FlockGpuError, match="Divide by zero") as info:
        gpu_ctx.filter_project(t, col(0) % col(1) == 0)
    assert info.value.code == -5
    assert gpu_ctx.filter_project(t, col(0) % 2 == 1).num_rows == 2      # the context stays usable


# ---- HashAggregateExec ---------------------------------------------------------------------------------------
def oracle_agg(b, mode, group, aggs):
    return oracle.hash_aggregate(b, mode, [(g, b.schema.names[g]) for g in group], [{"func": f, "col": c, "name": n} for f, c, n in aggs])


# AVG / SUM arguments are chosen so that every partial sum is an exactly representable integer (< 2^53): then the
# result does not depend on the order in which a GPU accumulates, and bit-exactness holds.
AGGS = [("count", -1, "COUNT(UInt8(1))"), ("sum", 1, "SUM(i64)"), ("min", 0, "MIN(i32)"), ("max", 2, "MAX(u64)"),
        ("avg", 0, "AVG(i32)"), ("max", 4, "MAX(ts)"), ("min", 3, "MIN(f64)"), ("sum", 0, "SUM(i32)")]


@pytest.mark.parametrize("group", [[6], [6, 0], [1], [5], [6, 5], [4, 2, 5]])
@pytest.mark.parametrize("n", [3, 5000, 300_000])
def test_hash_aggregate_single(gpu_ctx, group, n):
    b = mixed_batch(n, seed=n + len(group))
    t = gpu_ctx.import_batches([b])
    aggs = AGGS[:6] if len(group) < 3 else AGGS[:2]
    got = gpu_ctx.hash_aggregate(t, group, aggs, "single").to_arrow()
    want = pa.Table.from_batches([oracle_agg(b, "Single", group, aggs)])
    oracle.assert_tables_equal(got, want)


@pytest.mark.parametrize("shape", ["unique", "repeats", "random_unique"])
def test_distinct_wide_key_unique_probe(gpu_ctx, shape):
    """DISTINCT (Int32, Utf8): the 4-byte column is probed for repeats first (one streaming count); a column that never
    repeats makes the answer the input, one that does sends the rows through the row table -- same relation either way."""
    n = 400_000
    rng = np.random.default_rng(5)
    ids = np.arange(1000, 1000 + n, dtype=np.int32)
    if shape == "repeats":
        ids[rng.integers(0, n, 5000)] = ids[rng.integers(0, n, 5000)]          # some ids twice, mostly with another name; a few exact duplicates
    elif shape == "random_unique":
        ids = rng.permutation(n).astype(np.int32) * 7                               # unique but neither dense nor in time order
    names = np.array(["n%d" % (k % 97) for k in range(n)])
    if shape == "repeats":
        dup = rng.integers(1, n, 300)
        ids[dup], names[dup] = ids[dup - 1], names[dup - 1]
    b = rb(k=pa.array(ids), s=pa.array(names), pad=pa.array(np.zeros(n, np.int64)))
    got = gpu_ctx.hash_aggregate(gpu_ctx.import_batches([b]), [0, 1], [], "single").to_arrow()
    want = pa.Table.from_batches([b]).select(["k", "s"]).group_by(["k", "s"], use_threads=False).aggregate([])
    assert (got.num_rows < n) == (shape == "repeats")
    oracle.assert_tables_equal(got, want)


def test_hash_aggregate_partial_then_final(gpu_ctx):
    b = mixed_batch(200_000, seed=77)
    halves = [gpu_ctx.import_batches([b.slice(0, 90_000)]), gpu_ctx.import_batches([b.slice(90_000)])]
    aggs = [("count", -1, "c"), ("avg", 6, "a"), ("max", 0, "m"), ("sum", 0, "s")]
    partials = [gpu_ctx.hash_aggregate(h, [6, 5], aggs, "partial") for h in halves]
    p0 = partials[0].to_batch()
    assert p0.schema.names == ["k", "s", "c[count]", "a[count]", "a[sum]", "m[max]", "s[sum]"]          # aggregate.json naming
    merged = gpu_ctx.concat(partials)
    final_aggs = [("count", 2, "c"), ("avg", 3, "a"), ("max", 5, "m"), ("sum", 6, "s")]
    got = gpu_ctx.hash_aggregate(merged, [0, 1], final_aggs, "final_partitioned").to_arrow()
    want = pa.Table.from_batches([oracle_agg(b, "Single", [6, 5], aggs)])
    oracle.assert_tables_equal(got, want)


def test_global_aggregate_and_empty_input(gpu_ctx):
    b = mixed_batch(123_457, seed=5)
    t = gpu_ctx.import_batches([b])
    aggs = [("min", 1, "MIN"), ("avg", 0, "AVG"), ("count", 5, "COUNT"), ("max", 3, "MAXF")]
    got = gpu_ctx.hash_aggregate(t, [], aggs, "single").to_arrow()
    want = pa.Table.from_batches([oracle_agg(b, "Single", [], aggs)])
    oracle.assert_tables_equal(got, want)
    e = gpu_ctx.hash_aggregate(gpu_ctx.import_batches([b.slice(0, 0)]), [], aggs, "single").to_arrow()
    assert e.num_rows == 1 and e["COUNT"].to_pylist() == [0] and e["MIN"].to_pylist() == [None] and e["AVG"].to_pylist() == [None]
    assert gpu_ctx.hash_aggregate(gpu_ctx.import_batches([b.slice(0, 0)]), [6], aggs, "single").num_rows == 0


def test_aggregate_key_edge_cases(gpu_ctx):
    # the packed-key table reserves ~0 as its empty marker: a real key with that bit pattern must still group
    keys = np.array([-1, -1, 0, 5, -1, 0], np.int64)
    t = gpu_ctx.import_batches([rb(k=pa.array(keys), v=pa.array(np.arange(6, dtype=np.int64)))])
    got = gpu_ctx.hash_aggregate(t, [0], [("count", -1, "n"), ("sum", 1, "s")], "single").to_arrow()
    assert sorted(zip(*[got.column(i).to_pylist() for i in range(3)])) == [(-1, 3, 5), (0, 2, 7), (5, 1, 3)]
    big = np.full(400_000, -1, np.int64)
    big[::1000] = 7
    t = gpu_ctx.import_batches([rb(k=pa.array(big), v=pa.array(np.ones(400_000, np.int64)))])
    got = gpu_ctx.hash_aggregate(t, [0], [("count", -1, "n")], "single").to_arrow()
    assert sorted(zip(got["k"].to_pylist(), got["n"].to_pylist())) == [(-1, 399_600), (7, 400)]


@pytest.mark.parametrize("shape", ["hashed_i64", "dense_i32", "rows_utf8"])
def test_hash_aggregate_large_tables(gpu_ctx, shape):
    """Level-2 tables beyond 2.4 M slots take the 16 Ki-slot emit tiles (agg_emit_kernel<*, 64>): hashed packed keys,
    the direct-address table behind the shared-memory histogram, and the row-representative table."""
    rng = np.random.default_rng(5)
    if shape == "hashed_i64":
        n = 3_000_000
        b = rb(k=pa.array(rng.integers(0, 1 << 22, n) * 1_000_003), v=pa.array(rng.integers(0, 100, n)))
        aggs = [("count", -1, "n"), ("sum", 1, "s"), ("max", 1, "m")]
    elif shape == "dense_i32":
        n = 6_000_000
        b = rb(k=pa.array((np.arange(n) // 2 + rng.integers(0, 50, n)).astype(np.int32) + 1000))
        aggs = [("count", -1, "n")]
    else:
        n = 1_300_000
        ids = rng.permutation(n)
        b = rb(k=pa.array(["p%07d" % i for i in ids]), v=pa.array(rng.integers(0, 100, n)))
        aggs = [("count", -1, "n"), ("min", 1, "lo")]
    t = gpu_ctx.import_batches(nexgen.split_batches(b, 65536))
    got = gpu_ctx.hash_aggregate(t, [0], aggs, "single").to_arrow()
    want = pa.Table.from_batches([oracle_agg(b, "Single", [0], aggs)])
    assert got.num_rows > 1_000_000
    oracle.assert_tables_equal(got, want)


# ---- HashJoinExec -----------------------------------------------------------------------------------------------
@pytest.mark.parametrize("keys", [([0], [0]), ([5], [5]), ([6, 0], [6, 0]), ([6, 5], [6, 5]), ([1], [1])])
def test_hash_join_matches_oracle(gpu_ctx, keys):
    l = mixed_batch(20_000, seed=1)
    r = mixed_batch(30_000, seed=2)
    if keys[0] == [1]:
        r = r.set_column(1, r.schema.field(1), l["i64"].take(pa.array(np.random.default_rng(3).integers(0, 20_000, 30_000))))
    got = gpu_ctx.hash_join(gpu_ctx.import_batches([l]), gpu_ctx.import_batches([r]), *keys).to_arrow()
    want = pa.Table.from_batches([oracle.hash_join(l, r, *keys)])
    assert got.num_rows == want.num_rows and got.num_rows > 0
    oracle.assert_tables_equal(got, want, check_names=False)


def test_hash_join_edge_cases(gpu_ctx):
    l = rb(a=pa.array([1, 1, 2, 3], pa.int32()), x=pa.array(["p", "q", "r", "s"]))
    r = rb(b=pa.array([1, 1, 3, 4], pa.int32()), y=pa.array([10, 20, 30, 40], pa.int64()))
    got = gpu_ctx.hash_join(gpu_ctx.import_batches([l]), gpu_ctx.import_batches([r]), [0], [0]).to_arrow()
    assert got.schema.names == ["a", "x", "b", "y"]                      # left ++ right
    assert sorted(zip(*[got.column(i).to_pylist() for i in range(4)])) == [(1, "p", 1, 10), (1, "p", 1, 20), (1, "q", 1, 10), (1, "q", 1, 20), (3, "s", 3, 30)]
    empty = gpu_ctx.hash_join(gpu_ctx.import_batches([l.slice(0, 0)]), gpu_ctx.import_batches([r]), [0], [0])
    assert empty.num_rows == 0 and empty.num_columns == 4
    none = gpu_ctx.hash_join(gpu_ctx.import_batches([l]), gpu_ctx.import_batches([r.slice(3)]), [0], [0])
    assert none.num_rows == 0
    with pytest.raises(fb.FlockGpuError):
        gpu_ctx.hash_join(gpu_ctx.import_batches([l]), gpu_ctx.import_batches([r]), [0], [1])   # Int32 vs Int64 keys


@pytest.mark.parametrize("key_type", [pa.int32(), pa.int64(), pa.uint64()])
@pytest.mark.parametrize("one_row_side", ["left", "right"])
def test_hash_join_against_a_single_row(gpu_ctx, key_type, one_row_side):
    """NEXMark q5 / q7 shape: a relation joined with a one-row global aggregate -- an equality filter on the GPU.
    Output columns stay left ++ right, matching probe rows keep their order."""
    n = 70_003
    rng = np.random.default_rng(11)
    k = rng.integers(0, 50, n)
    big = rb(k=pa.array(k, key_type), v=pa.array(np.arange(n, dtype=np.int64)), s=pa.array(["r%d" % (x % 13) for x in range(n)]))
    for wanted in (7, 1000):                                              # present many times / absent
        one = rb(m=pa.array([wanted], key_type), tag=pa.array(["only"]))
        l, r = (one, big) if one_row_side == "left" else (big, one)
        lk, rk = [0], [0]
        got = gpu_ctx.hash_join(gpu_ctx.import_batches([l]), gpu_ctx.import_batches([r]), lk, rk).to_arrow()
        want = pa.Table.from_batches([oracle.hash_join(l, r, lk, rk)])
        assert got.num_rows == int((k == wanted).sum()) == want.num_rows
        assert got.schema.names == l.schema.names + r.schema.names
        oracle.assert_tables_equal(got, want, check_names=False)
        if got.num_rows:
            assert got["v"].to_pylist() == np.nonzero(k == wanted)[0].tolist()      # probe order preserved


# ---- COUNT / DISTINCT by one 4-byte key: the direct-address table as a deferred relation ------------------------------
@pytest.mark.parametrize("shape", ["dense", "dense_distinct", "sparse", "two_clusters"])
def test_dense_count_table_and_its_fallback(gpu_ctx, shape):
    """Large inputs with one Int32 group column aggregate into a direct-address table that stays in table form until a
    consumer needs rows (hash_agg.cu: DeferredTable).  Keys outside the sampled range (sparse / clustered far apart)
    make the first consumer start over on the general path: the results must not differ either way."""
    rng = np.random.default_rng(11)
    n = 700_000
    if shape in ("dense", "dense_distinct"):
        k = (np.arange(n) // 13 + rng.integers(0, 40, n) + 1000).astype(np.int32)      # advancing ids, like NEXMark's
    elif shape == "sparse":
        k = rng.integers(-(1 << 31), 1 << 31, n).astype(np.int32)
    else:
        k = np.where(rng.integers(0, 2, n) == 0, rng.integers(0, 5000, n), rng.integers(1 << 30, (1 << 30) + 5000, n)).astype(np.int32)
    b = rb(k=pa.array(k), v=pa.array(rng.integers(0, 1000, n)))
    t = gpu_ctx.import_batches([b])
    aggs = [] if shape == "dense_distinct" else [("count", -1, "n")]
    for mode in ("single", "partial"):
        got = gpu_ctx.hash_aggregate(t, [0], aggs, mode)
        want = pa.Table.from_batches([oracle_agg(b, mode.capitalize(), [0], aggs)])
        assert got.num_rows == want.num_rows == np.unique(k).size
        oracle.assert_tables_equal(got.to_arrow(), want)
    # the deferred relation under every kind of consumer
    d = gpu_ctx.hash_aggregate(t, [0], [("count", -1, "n")], "single")
    mx = gpu_ctx.hash_aggregate(d, [], [("max", 1, "m")], "single")                     # fast path: MAX over the table
    counts = np.unique(k, return_counts=True)
    assert mx.to_arrow()["m"].to_pylist() == [int(counts[1].max())]
    top = gpu_ctx.hash_join(d, mx, [1], [0]).to_arrow()                                 # fast path: count = max
    assert sorted(top["k"].to_pylist()) == sorted(counts[0][counts[1] == counts[1].max()].tolist())
    assert set(top["n"].to_pylist()) == {int(counts[1].max())} and top.schema.names == ["k", "n", "m"]
    flipped = gpu_ctx.hash_join(mx, d, [0], [1]).to_arrow()
    assert flipped.schema.names == ["m", "k", "n"] and sorted(flipped["k"].to_pylist()) == sorted(top["k"].to_pylist())
    d2 = gpu_ctx.hash_aggregate(t, [0], [("count", -1, "n")], "single")
    kept = gpu_ctx.filter_project(d2, col(1) > lit(13, "uint64")).to_arrow()                           # generic consumer: materialises first
    assert sorted(kept["k"].to_pylist()) == sorted(counts[0][counts[1] > 13].tolist())
    renamed = gpu_ctx.filter_project(gpu_ctx.hash_aggregate(t, [0], [("count", -1, "n")], "single"), None, [col(1), col(0)], ["num", "key"])
    back = renamed.to_arrow()
    assert back.schema.names == ["num", "key"]
    order = np.argsort(back["key"].to_numpy(), kind="stable")
    assert np.array_equal(back["key"].to_numpy()[order], counts[0]) and np.array_equal(back["num"].to_numpy()[order], counts[1].astype(np.uint64))


# ---- SortExec / WindowAggExec(ROW_NUMBER) / GlobalLimitExec ---------------------------------------------------------------
@pytest.mark.parametrize("n", [0, 1, 2, 1000, 200_003])
def test_sort_row_number_limit(gpu_ctx, n):
    rng = np.random.default_rng(n + 1)
    b = rb(g=pa.array(rng.integers(-5, 40, n).astype(np.int32)), t=pa.array(rng.integers(1_436_918_400_000, 1_436_918_400_500, n), pa.timestamp("ms")),
           p=pa.array(rng.integers(0, 1 << 40, n)), u=pa.array(rng.integers(0, 1 << 63, n).astype(np.uint64)), f=pa.array(rng.normal(0, 100, n)),
           k=pa.array(rng.integers(0, 3, n).astype(np.int32)))
    t = gpu_ctx.import_batches([b])
    for cols, desc in (([0], [False]), ([0, 1], [False, True]), ([4], [True]), ([5, 3], [True, False]), ([1, 2], [True, True])):
        got = gpu_ctx.sort(t, cols, desc).to_batch()
        spec = [{"expr": plans.column(b.schema.names[c], c), "options": {"descending": d, "nulls_first": d}} for c, d in zip(cols, desc)]
        want = oracle.sort_batch(b, spec)
        assert got.equals(want), (cols, desc)
    # ROW_NUMBER over the sorted relation, then the filter / limit q6 puts on top
    s = gpu_ctx.sort(t, [0, 2], [False, True])
    w = gpu_ctx.row_number(s, [0], "rn")
    got = w.to_batch()
    want = oracle.window_batch(oracle.sort_batch(b, [{"expr": plans.column("g", 0), "options": {"descending": False}},
                                                     {"expr": plans.column("p", 2), "options": {"descending": True}}]),
                               [{"fun": "RowNumber", "name": "rn", "partition_by": [plans.column("g", 0)], "order_by": []}])
    assert got.schema.names == ["rn", "g", "t", "p", "u", "f", "k"] and got.equals(want)
    both = gpu_ctx.row_number(gpu_ctx.sort(t, [0, 5, 1], [False, False, False]), [0, 5], "rn2").to_batch()
    keys = list(zip(both["g"].to_pylist(), both["k"].to_pylist()))
    seen = {}
    for key, r in zip(keys, both["rn2"].to_pylist()):
        seen[key] = seen.get(key, 0) + 1
        assert r == seen[key]
    for lim in (0, 1, 7, n, n + 5):
        assert gpu_ctx.limit(s, lim).to_batch().equals(s.to_batch().slice(0, min(lim, n)))
    names = rb(s=pa.array(["b", "a", "c"]), v=pa.array([1, 2, 3]))
    assert gpu_ctx.limit(gpu_ctx.import_batches([names]), 2).to_batch().equals(names.slice(0, 2))
    # Utf8 sort keys and tie-breakers: byte order, shorter first, embedded NUL, multi-byte characters, > 8 bytes
    words = ["", "a", "a\0", "ab", "abcdefgh", "abcdefghi", "abcdefgh\0", "b", "émile", "日本", "zzzzzzzzzzzzzzzzzzzzzzzzzq", "zzzzzzzzzzzzzzzzzzzzzzzzzp"]
    m = max(n, 1)
    sb = rb(s=pa.array([words[i] for i in rng.integers(0, len(words), m)]), v=pa.array(rng.integers(0, 4, m)))
    st = gpu_ctx.import_batches([sb])
    for cols, desc in (([0], [False]), ([0], [True]), ([1], [False]), ([1, 0], [True, True])):
        spec = [{"expr": plans.column(sb.schema.names[c], c), "options": {"descending": d}} for c, d in zip(cols, desc)]
        assert gpu_ctx.sort(st, cols, desc).to_batch().equals(oracle.sort_batch(sb, spec)), (cols, desc)


# ---- RepartitionExec: Hash -------------------------------------------------------------------------------------------
@pytest.mark.parametrize("keys", [[0], [5], [6, 0], [1]])
@pytest.mark.parametrize("n_parts", [2, 8])
def test_hash_partition(gpu_ctx, keys, n_parts):
    b = mixed_batch(100_001, seed=42)
    parts = [p.to_batch() for p in gpu_ctx.hash_partition(gpu_ctx.import_batches([b]), keys, n_parts)]
    assert len(parts) == n_parts and sum(p.num_rows for p in parts) == b.num_rows
    if all(pa.types.is_integer(b.schema.field(k).type) for k in keys):
        pid = sharding.partition_ids(b, keys, n_parts)                   # the numpy model of the device hash
        for q, p in enumerate(parts):
            assert p.equals(b.filter(pa.array(pid == q)))                # membership AND input order inside a partition
    else:
        seen = {}
        for q, p in enumerate(parts):
            for key in zip(*[p.column(k).to_pylist() for k in keys]):
                assert seen.setdefault(key, q) == q
        oracle.assert_tables_equal(pa.Table.from_batches(parts), pa.Table.from_batches([b]))


@pytest.mark.parametrize("n_parts", [3, 33, 255])
def test_hash_partition_many_ways(gpu_ctx, n_parts):
    """One multi-way pass whatever n is; a fixed-width key column next to a Utf8 one routes alone (partition.cu:
    routing_columns), so the numpy model over [i32] must reproduce membership and order for keys [i32, s]."""
    b = mixed_batch(300_007, seed=n_parts)
    parts = [p.to_batch() for p in gpu_ctx.hash_partition(gpu_ctx.import_batches([b]), [0, 5], n_parts)]
    pid = sharding.partition_ids(b, [0], n_parts)
    assert len(parts) == n_parts
    for q, p in enumerate(parts):
        assert p.equals(b.filter(pa.array(pid == q))), q
    # views of one partition-ordered buffer set are ordinary tables: they filter, aggregate and concatenate
    tabs = gpu_ctx.hash_partition(gpu_ctx.import_batches([b]), [6], 4)
    back = gpu_ctx.concat(tabs).to_arrow()
    oracle.assert_tables_equal(back, pa.Table.from_batches([b]))
    pid6 = sharding.partition_ids(b, [6], 4)
    for q, t in enumerate(tabs):
        got = gpu_ctx.filter_project(t, col(0).cast("int64") % 7 == 0).to_batch()
        assert got.equals(oracle_filter(b.filter(pa.array(pid6 == q)), col(0).cast("int64") % 7 == 0))


def test_hash_partition_long_strings_and_empty_partitions(gpu_ctx):
    """Tiles whose strings exceed the 40 KB shared staging take the direct byte path; destinations nobody routes to
    come back as empty relations with the right schema."""
    rng = np.random.default_rng(5)
    n = 20_000
    lens = rng.integers(0, 600, n)
    s = pa.array(["".join(chr(97 + (i + j) % 26) for j in range(l)) if l < 40 else ("q%d" % i) * (l // 4) for i, l in enumerate(lens)])
    b = rb(k=pa.array(np.full(n, 7, np.int32)), s=s, v=pa.array(np.arange(n, dtype=np.int64)))
    parts = [p.to_batch() for p in gpu_ctx.hash_partition(gpu_ctx.import_batches([b]), [0], 5)]
    home = int(sharding.partition_ids(b.slice(0, 1), [0], 5)[0])
    for q, p in enumerate(parts):
        assert p.schema.names == ["k", "s", "v"]
        assert p.equals(b) if q == home else p.num_rows == 0


def test_reversed_literal_division_on_a_ragged_tile(gpu_ctx):
    """lit / (col * k) and lit % (col * k): the padded rows of a tile must not raise a divide-by-zero (they carry 0)."""
    b = rb(x=pa.array([3, 5, -2], pa.int64()))
    t = gpu_ctx.import_batches([b])
    got = gpu_ctx.filter_project(t, None, [lit(10) / (col(0) * 2), lit(7) % (col(0) * 3)], ["d", "m"]).to_batch()
    assert got["d"].to_pylist() == [1, 1, -2] and got["m"].to_pylist() == [7, 7, 1]
    kept = gpu_ctx.filter_project(t, (lit(10) / (col(0) * 2)) == 1).to_batch()
    assert kept["x"].to_pylist() == [3, 5]
    with pytest.raises(fb.FlockGpuError) as info:
        gpu_ctx.filter_project(gpu_ctx.import_batches([rb(x=pa.array([3, 0, 1], pa.int64()))]), None, [lit(10) / (col(0) * 2)], ["d"]).to_batch()
    assert info.value.code == -5


# ---- the three grid-wide prefix protocols of compact.cuh give identical results ------------------------------------
@pytest.mark.parametrize("mode", [0, 1])
def test_prefix_protocols_agree(gpu_ctx, mode):
    """compact_mode 0 = automatic (single wave: arrival counter + self-validating count words; decoupled look-back
    beyond one wave), 1 = decoupled look-back always."""
    gpu_ctx.set_option("compact_mode", mode)
    try:
        for n in (5_000, 250_000, 1_500_000):
            rng = np.random.default_rng(n)
            b = rb(k=pa.array(rng.integers(0, 1 << 20, n).astype(np.int32)), v=pa.array(rng.integers(0, 1000, n)),
                   s=pa.array([("w%d" % (x % 97)) * (x % 3) for x in rng.integers(0, 1 << 30, n)]))
            t = gpu_ctx.import_batches([b])
            # filter: vectorised functor (col 0) and generic interpreter with a Utf8 pass-through (selection vector + gather)
            for pred in (col(0).cast("int64") % 7 == 0, (col(1) < 300) | (col(2) == "w5")):
                got = gpu_ctx.filter_project(t, pred).to_batch()
                assert got.equals(oracle_filter(b, pred))
            # aggregate emit (hashed table) and join count scan
            got = gpu_ctx.hash_aggregate(t, [0], [("count", -1, "n"), ("sum", 1, "s")], "single").to_arrow()
            want = pa.Table.from_batches([oracle_agg(b, "Single", [0], [("count", -1, "n"), ("sum", 1, "s")])])
            oracle.assert_tables_equal(got, want)
            if n <= 250_000:
                r = b.slice(0, n // 3)
                got = gpu_ctx.hash_join(gpu_ctx.import_batches([r]), t, [0], [0]).to_arrow()
                want = pa.Table.from_batches([oracle.hash_join(r, b, [0], [0])])
                oracle.assert_tables_equal(got, want, check_names=False)
            parts = gpu_ctx.hash_partition(t, [0], 3)
            pid = sharding.partition_ids(b, [0], 3)
            for q, p in enumerate(parts):
                assert p.to_batch().equals(b.filter(pa.array(pid == q)))
    finally:
        gpu_ctx.set_option("compact_mode", 0)
