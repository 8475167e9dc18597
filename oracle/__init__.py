"""CPU oracle: a restatement of the DataFusion-6 physical operators Flock's hot path runs.

TEST INFRASTRUCTURE ONLY -- imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
``--impl reference`` legs, never by flock_b200 (the product fails loudly without its CUDA library).

PARITY UNPINNED for NEXMark q1-q8 outputs: no reference test asserts them (every
flock/src/datasource/nexmark/queries/qN.rs only println!s) and the reference binary cannot be built
here (no Rust toolchain; operators live in an un-vendored DataFusion fork, flock/Cargo.toml:21).  The
oracle is pinned instead against the reference's toy goldens (flock/src/runtime/context.rs:428-592,
flock/src/launcher/local.rs:169-234) and an independent Arrow C++ implementation (oracle/acero_ref.py).

``execute_plan`` interprets the reference's serde-JSON physical plan (the same JSON the GPU executor
takes) the way DataFusion 6 executes it: ``target_partitions`` streams, RoundRobinBatch / Hash
repartitioning between them, Partial -> FinalPartitioned aggregation, partition-wise hash joins,
CoalesceBatches(4096).  Operator arithmetic is in oracle/oracle.cc (C++), materialisation (take /
filter / concat) uses the Arrow C++ kernels through pyarrow.
"""
from __future__ import annotations

import ctypes as C
import json
import subprocess
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

import numpy as np
import pyarrow as pa

HERE = Path(__file__).resolve().parent
LIB_PATH = HERE / "liboracle.so"

T_BOOL, T_I32, T_I64, T_U64, T_F64, T_TS, T_UTF8, T_U32 = range(8)
OP = dict(COLUMN=1, LIT_I64=2, LIT_F64=3, LIT_UTF8=4, CAST=5, ADD=10, SUB=11, MUL=12, DIV=13, MOD=14,
          EQ=20, NE=21, LT=22, LE=23, GT=24, GE=25, AND=30, OR=31, NOT=32)
A_COUNT, A_SUM_I, A_SUM_F, A_MIN_I, A_MAX_I, A_MIN_U, A_MAX_U, A_MIN_F, A_MAX_F = range(9)


def build(force: bool = False) -> Path:
    src = HERE / "oracle.cc"
    if force or not LIB_PATH.exists() or LIB_PATH.stat().st_mtime < src.stat().st_mtime:
        subprocess.run(["make", "-C", str(HERE), "-s", "liboracle.so"], check=True)
    return LIB_PATH


class OCol(C.Structure):
    _fields_ = [("dtype", C.c_int32), ("pad", C.c_int32), ("len", C.c_int64), ("data", C.c_void_p), ("offsets", C.c_void_p)]


class OTok(C.Structure):
    _fields_ = [("op", C.c_int32), ("dtype", C.c_int32), ("col", C.c_int32), ("str_len", C.c_int32),
                ("i64", C.c_int64), ("f64", C.c_double), ("str", C.c_char_p)]


class OAcc(C.Structure):
    _fields_ = [("kind", C.c_int32), ("col", C.c_int32)]


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(str(LIB_PATH))
        _lib.orc_eval_predicate.restype = C.c_int
        _lib.orc_eval_value.restype = C.c_int
        _lib.orc_infer_dtype.restype = C.c_int32
        _lib.orc_mask_to_indices.restype = C.c_int64
        _lib.orc_partition_ids.restype = C.c_int
        _lib.orc_group_by.restype = C.c_int64
        _lib.orc_hash_join.restype = C.c_int64
        _lib.orc_q2_batch.restype = C.c_int64
        _lib.orc_q2_collect.restype = C.c_int64
    return _lib


# ------------------------------------------------------------------------------------------------
# pyarrow <-> OCol
# ------------------------------------------------------------------------------------------------
def type_code(t: pa.DataType) -> int:
    if pa.types.is_int32(t): return T_I32
    if pa.types.is_uint32(t): return T_U32
    if pa.types.is_int64(t): return T_I64
    if pa.types.is_uint64(t): return T_U64
    if pa.types.is_float64(t): return T_F64
    if pa.types.is_timestamp(t): return T_TS
    if pa.types.is_string(t): return T_UTF8
    raise TypeError(f"oracle: unsupported type {t}")


_NP = {T_I32: np.int32, T_U32: np.uint32, T_I64: np.int64, T_U64: np.uint64, T_F64: np.float64, T_TS: np.int64}


class Cols:
    """The columns of one RecordBatch as an OCol array (keeps the backing arrays alive)."""

    def __init__(self, batch: pa.RecordBatch):
        self.keep = []
        self.n = batch.num_rows
        self.arr = (OCol * max(batch.num_columns, 1))()
        for i, col in enumerate(batch.columns):
            if col.null_count:
                raise ValueError("oracle: nulls are not supported (the NEXMark schemas are non-nullable)")
            tc = type_code(col.type)
            o = self.arr[i]
            o.dtype, o.len = tc, len(col)
            bufs = col.buffers()
            if tc == T_UTF8:
                offs = np.frombuffer(bufs[1], dtype=np.int32, count=len(col) + 1 + col.offset)[col.offset:]
                offs = np.ascontiguousarray(offs)
                self.keep += [offs, bufs[2]]
                o.offsets = offs.ctypes.data
                o.data = bufs[2].address if bufs[2] is not None else 0
            else:
                w = np.dtype(_NP[tc]).itemsize
                o.data = bufs[1].address + col.offset * w if len(col) else 0
                self.keep.append(bufs[1])
        self.keep.append(batch)


def _arrow_type(tc: int, like: pa.DataType | None = None) -> pa.DataType:
    if tc == T_TS:
        return like if like is not None and pa.types.is_timestamp(like) else pa.timestamp("ms")
    return {T_I32: pa.int32(), T_U32: pa.uint32(), T_I64: pa.int64(), T_U64: pa.uint64(), T_F64: pa.float64()}[tc]


# ------------------------------------------------------------------------------------------------
# expressions: reference JSON -> postfix tokens
# ------------------------------------------------------------------------------------------------
_BINOP = {"Eq": "EQ", "NotEq": "NE", "Lt": "LT", "LtEq": "LE", "Gt": "GT", "GtEq": "GE", "Plus": "ADD", "Minus": "SUB",
          "Multiply": "MUL", "Divide": "DIV", "Modulo": "MOD", "Modulus": "MOD", "And": "AND", "Or": "OR"}


def _dtype_code_json(t) -> int:
    if isinstance(t, dict):
        if "Timestamp" in t:
            return T_TS
        raise TypeError(f"oracle: unsupported cast type {t}")
    return {"Int32": T_I32, "UInt32": T_U32, "Int64": T_I64, "UInt64": T_U64, "Float64": T_F64, "Utf8": T_UTF8}[t]


def _resolve(schema_names: list[str], e: dict) -> int:
    idx = e.get("index")
    if idx is not None and 0 <= idx < len(schema_names) and schema_names[idx] == e["name"]:
        return idx
    if e["name"] in schema_names:
        return schema_names.index(e["name"])
    if idx is not None:
        return idx
    raise KeyError(e["name"])


def expr_tokens(e: dict, names: list[str]) -> list[tuple]:
    k = e["physical_expr"]
    if k == "column":
        return [(OP["COLUMN"], 0, _resolve(names, e), 0, 0.0, None)]
    if k == "literal":
        (ty, v), = e["value"].items()
        if v is None:
            raise ValueError("oracle: NULL literal")
        if ty in ("Float64", "Float32"):
            return [(OP["LIT_F64"], T_F64, 0, 0, float(v), None)]
        if ty in ("Utf8", "LargeUtf8"):
            return [(OP["LIT_UTF8"], T_UTF8, 0, 0, 0.0, v.encode())]
        dt = {"Int32": T_I32, "UInt32": T_U32, "UInt64": T_U64}.get(ty, T_TS if ty.startswith("Timestamp") else T_I64)
        iv = int(v)
        return [(OP["LIT_I64"], dt, 0, iv if iv < (1 << 63) else iv - (1 << 64), 0.0, None)]
    if k in ("cast_expr", "try_cast_expr"):
        return expr_tokens(e["expr"], names) + [(OP["CAST"], _dtype_code_json(e["cast_type"]), 0, 0, 0.0, None)]
    if k == "binary_expr":
        return expr_tokens(e["left"], names) + expr_tokens(e["right"], names) + [(OP[_BINOP[e["op"]]], 0, 0, 0, 0.0, None)]
    if k in ("not_expr", "not"):
        return expr_tokens(e.get("arg", e.get("expr")), names) + [(OP["NOT"], 0, 0, 0, 0.0, None)]
    raise ValueError(f"oracle: unsupported physical expression {k}")


def _ctoks(tokens: list[tuple]):
    arr = (OTok * len(tokens))()
    keep = []
    for i, (op, dt, c, i64, f64, s) in enumerate(tokens):
        t = arr[i]
        t.op, t.dtype, t.col, t.i64, t.f64 = op, dt, c, i64, f64
        if s is not None:
            keep.append(s)
            t.str, t.str_len = s, len(s)
    return arr, keep


class OracleError(RuntimeError):
    pass


def _check(rc: int, what: str):
    if rc == -5:
        raise OracleError("Divide by zero")
    if rc:
        raise OracleError(f"oracle: {what} failed with code {rc}")


def eval_predicate(batch: pa.RecordBatch, expr: dict) -> np.ndarray:
    cols = Cols(batch)
    toks, keep = _ctoks(expr_tokens(expr, batch.schema.names))
    mask = np.zeros(batch.num_rows, np.uint8)
    _check(lib().orc_eval_predicate(cols.arr, batch.num_columns, C.c_int64(batch.num_rows), toks, len(toks), mask.ctypes.data_as(C.c_void_p)),
           "predicate")
    return mask.astype(bool)


def eval_value(batch: pa.RecordBatch, expr: dict) -> pa.Array:
    if expr["physical_expr"] == "column":
        return batch.column(_resolve(batch.schema.names, expr))      # zero-copy, like the Arc clone in ProjectionExec
    cols = Cols(batch)
    toks, keep = _ctoks(expr_tokens(expr, batch.schema.names))
    tc = lib().orc_infer_dtype(cols.arr, batch.num_columns, toks, len(toks))
    if tc < 0:
        raise OracleError("oracle: cannot type the expression")
    out = np.zeros(batch.num_rows, _NP[tc])
    got = C.c_int32()
    _check(lib().orc_eval_value(cols.arr, batch.num_columns, C.c_int64(batch.num_rows), toks, len(toks), out.ctypes.data_as(C.c_void_p), C.byref(got)),
           "value expression")
    like = None
    if tc == T_TS:
        like = next((f.type for f in batch.schema if pa.types.is_timestamp(f.type)), None)
    arr = pa.array(out)
    return arr.cast(_arrow_type(tc, like)) if tc == T_TS else arr


# ------------------------------------------------------------------------------------------------
# operators
# ------------------------------------------------------------------------------------------------
def _concat(batches: list[pa.RecordBatch], schema: pa.Schema) -> pa.RecordBatch:
    if not batches:
        return pa.RecordBatch.from_arrays([pa.array([], f.type) for f in schema], schema=schema)
    if len(batches) == 1:
        return batches[0]
    t = pa.Table.from_batches(batches).combine_chunks()
    return t.to_batches()[0] if t.num_rows else batches[0].slice(0, 0)


def filter_batch(batch: pa.RecordBatch, predicate: dict) -> pa.RecordBatch:
    return batch.filter(pa.array(eval_predicate(batch, predicate)))


def project_batch(batch: pa.RecordBatch, exprs: list) -> pa.RecordBatch:
    arrays = [eval_value(batch, e) for e, _ in exprs]
    fields = []
    for (e, name), a in zip(exprs, arrays):
        nullable = True
        if e["physical_expr"] == "column":
            nullable = batch.schema.field(_resolve(batch.schema.names, e)).nullable
        fields.append(pa.field(name, a.type, nullable))
    return pa.RecordBatch.from_arrays(arrays, schema=pa.schema(fields, metadata=batch.schema.metadata))


def partition_ids(batch: pa.RecordBatch, key_cols: list[int], n_parts: int) -> np.ndarray:
    cols = Cols(batch)
    keys = (C.c_int32 * len(key_cols))(*key_cols)
    pid = np.zeros(batch.num_rows, np.int32)
    _check(lib().orc_partition_ids(cols.arr, keys, len(key_cols), C.c_int64(batch.num_rows), n_parts, pid.ctypes.data_as(C.c_void_p)), "partition")
    return pid


_AGG_STATE = {"count": ["count"], "sum": ["sum"], "min": ["min"], "max": ["max"], "avg": ["count", "sum"]}


def _minmax_kind(t: pa.DataType, is_min: bool) -> int:
    if pa.types.is_float64(t):
        return A_MIN_F if is_min else A_MAX_F
    if pa.types.is_uint64(t) or pa.types.is_uint32(t):
        return A_MIN_U if is_min else A_MAX_U
    return A_MIN_I if is_min else A_MAX_I


def _sortable(arr: pa.Array) -> np.ndarray:
    """Dense rank of every value in the column's ascending order (Utf8: byte order, like Arrow) as int64: negating it
    gives the descending order without overflow for any type."""
    if pa.types.is_string(arr.type):
        vals = np.array([v.encode("utf-8") for v in arr.to_pylist()], dtype=object)
    elif pa.types.is_timestamp(arr.type):
        vals = arr.cast(pa.int64()).to_numpy(zero_copy_only=False)
    else:
        vals = arr.to_numpy(zero_copy_only=False)
    _, inv = np.unique(vals, return_inverse=True)
    return inv.astype(np.int64)


def sort_batch(batch: pa.RecordBatch, sort_exprs: list[dict]) -> pa.RecordBatch:
    """Lexicographic sort by [{expr: column, options: {descending, nulls_first}}]; ties by the other columns ascending."""
    if batch.num_rows == 0:
        return batch
    names = batch.schema.names
    used, keys = [], []
    for se in sort_exprs:
        i = _resolve(names, se["expr"])
        used.append(i)
        k = _sortable(batch.column(i))
        keys.append(-k if se.get("options", {}).get("descending", False) else k)
    for i in range(batch.num_columns):                     # deterministic tie-break
        if i not in used:
            keys.append(_sortable(batch.column(i)))
    order = np.lexsort(tuple(reversed(keys)))              # np.lexsort: LAST key is primary
    return batch.take(pa.array(order))


def window_batch(batch: pa.RecordBatch, window_exprs: list[dict]) -> pa.RecordBatch:
    """ROW_NUMBER() OVER (PARTITION BY p.. ORDER BY o..) on an input already sorted by (p.., o..)."""
    names = batch.schema.names
    n = batch.num_rows
    cols, fields = [], []
    for w in window_exprs:
        if w.get("fun") != "RowNumber":
            raise OracleError(f"oracle: window function {w.get('fun')} is not restated")
        part = [_resolve(names, e) for e in w.get("partition_by", [])]
        start = np.ones(n, bool)
        if n:
            start[1:] = False
            for i in part:
                k = _sortable(batch.column(i))
                start[1:] |= k[1:] != k[:-1]
        idx = np.arange(n, dtype=np.int64)
        first = np.maximum.accumulate(np.where(start, idx, 0))
        cols.append(pa.array((idx - first + 1).astype(np.uint64)))
        fields.append(pa.field(w["name"], pa.uint64(), True))
    return pa.RecordBatch.from_arrays(cols + list(batch.columns), schema=pa.schema(fields + list(batch.schema), metadata=batch.schema.metadata))


def hash_aggregate(batch: pa.RecordBatch, mode: str, group: list[tuple[int, str]], aggrs: list[dict]) -> pa.RecordBatch:
    """One partition of HashAggregateExec.  aggrs: [{func, col (or first state col), name}]."""
    final = mode in ("Final", "FinalPartitioned")
    accs, outs = [], []   # outs: (name, kind, a0, a1, arrow type)
    for a in aggrs:
        f, col, name = a["func"], a["col"], a["name"]
        t = batch.schema.field(col).type if col is not None and col >= 0 else None
        if f == "count":
            accs.append((A_SUM_I, col) if final else (A_COUNT, -1))
            outs.append((name if final or mode == "Single" else name + "[count]", "raw", len(accs) - 1, None, pa.uint64()))
        elif f == "sum":
            if pa.types.is_float64(t):
                accs.append((A_SUM_F, col)); ot = pa.float64()
            elif pa.types.is_uint64(t) or pa.types.is_uint32(t):
                accs.append((A_SUM_I, col)); ot = pa.uint64()
            else:
                accs.append((A_SUM_I, col)); ot = pa.int64()
            outs.append((name if final or mode == "Single" else name + "[sum]", "raw", len(accs) - 1, None, ot))
        elif f in ("min", "max"):
            accs.append((_minmax_kind(t, f == "min"), col))
            outs.append((name if final or mode == "Single" else f"{name}[{f}]", "raw", len(accs) - 1, None, t))
        elif f == "avg":
            if final:
                accs.append((A_SUM_I, col)); accs.append((A_SUM_F, col + 1))
            else:
                accs.append((A_COUNT, -1)); accs.append((A_SUM_F, col))
            if final or mode == "Single":
                outs.append((name, "avg", len(accs) - 2, len(accs) - 1, pa.float64()))
            else:
                outs.append((name + "[count]", "raw", len(accs) - 2, None, pa.uint64()))
                outs.append((name + "[sum]", "raw", len(accs) - 1, None, pa.float64()))
        else:
            raise ValueError(f)
    if mode == "Single":   # Partial + Final in one step: COUNT counts rows
        pass
    n = batch.num_rows
    cols = Cols(batch)
    keys = (C.c_int32 * max(len(group), 1))(*[g for g, _ in group])
    cacc = (OAcc * max(len(accs), 1))()
    for i, (k, c) in enumerate(accs):
        cacc[i].kind, cacc[i].col = k, (c if c is not None else -1)
    first = np.zeros(max(n, 1), np.int64)
    state = np.zeros(max(n, 1) * max(len(accs), 1), np.uint64)
    ng = lib().orc_group_by(cols.arr, keys, len(group), C.c_int64(n), cacc, len(accs), first.ctypes.data_as(C.c_void_p), state.ctypes.data_as(C.c_void_p))
    arrays, fields = [], []
    if not group and ng == 0:
        # global aggregate over empty input: one row, COUNT = 0, everything else NULL (Appendix C.7)
        for name, kind, a0, a1, ot in outs:
            is_count = kind == "raw" and accs[a0][0] in (A_COUNT,) or (final and kind == "raw" and ot == pa.uint64() and accs[a0][0] == A_SUM_I)
            arrays.append(pa.array([0], ot) if is_count else pa.array([None], ot))
            fields.append(pa.field(name, ot, True))
        return pa.RecordBatch.from_arrays(arrays, schema=pa.schema(fields))
    idx = pa.array(first[:ng])
    for g, name in group:
        arrays.append(batch.column(g).take(idx))
        fields.append(pa.field(name, batch.schema.field(g).type, batch.schema.field(g).nullable))
    st = state.reshape(max(len(accs), 1), max(n, 1))
    for name, kind, a0, a1, ot in outs:
        if kind == "avg":
            cnt = st[a0, :ng].astype(np.float64)
            sm = st[a1, :ng].view(np.float64)
            arrays.append(pa.array(sm / cnt))
        else:
            raw = st[a0, :ng]
            if pa.types.is_float64(ot):
                arrays.append(pa.array(raw.view(np.float64)))
            elif pa.types.is_uint64(ot):
                arrays.append(pa.array(raw))
            elif pa.types.is_uint32(ot):
                arrays.append(pa.array(raw.astype(np.uint32)))
            elif pa.types.is_int32(ot):
                arrays.append(pa.array(raw.view(np.int64).astype(np.int32)))
            elif pa.types.is_timestamp(ot):
                arrays.append(pa.array(raw.view(np.int64)).cast(ot))
            else:
                arrays.append(pa.array(raw.view(np.int64)))
        fields.append(pa.field(name, ot, True))
    return pa.RecordBatch.from_arrays(arrays, schema=pa.schema(fields, metadata=batch.schema.metadata))


def hash_join(left: pa.RecordBatch, right: pa.RecordBatch, lkeys: list[int], rkeys: list[int]) -> pa.RecordBatch:
    lc, rc = Cols(left), Cols(right)
    lk = (C.c_int32 * len(lkeys))(*lkeys)
    rk = (C.c_int32 * len(rkeys))(*rkeys)
    args = (lc.arr, lk, C.c_int64(left.num_rows), rc.arr, rk, C.c_int64(right.num_rows), len(lkeys))
    m = lib().orc_hash_join(*args, None, None, C.c_int64(0))
    li, ri = np.zeros(max(m, 1), np.int64), np.zeros(max(m, 1), np.int64)
    lib().orc_hash_join(*args, li.ctypes.data_as(C.c_void_p), ri.ctypes.data_as(C.c_void_p), C.c_int64(m))
    li, ri = pa.array(li[:m]), pa.array(ri[:m])
    arrays = [c.take(li) for c in left.columns] + [c.take(ri) for c in right.columns]
    fields = list(left.schema) + list(right.schema)
    return pa.RecordBatch.from_arrays(arrays, schema=pa.schema(fields))


# ------------------------------------------------------------------------------------------------
# plan interpreter
# ------------------------------------------------------------------------------------------------
Partitions = list  # list[list[pa.RecordBatch]]


def q2_collect(batches: list[pa.RecordBatch], n_partitions: int, n_threads: int, modulus: int = 123, rhs: int = 0, repeat: int = 1):
    """NEXMark q2 over a relation by the native partition-parallel pipeline (orc_q2_collect: native threads, no Python per batch).
    Returns (result table, seconds per call as a list of `repeat` wall times)."""
    import time
    n = len(batches)
    au = (C.c_void_p * n)(*[b.column("auction").buffers()[1].address + 4 * b.column("auction").offset for b in batches])
    pr = (C.c_void_p * n)(*[b.column("price").buffers()[1].address + 4 * b.column("price").offset for b in batches])
    rows = (C.c_int64 * n)(*[b.num_rows for b in batches])
    total = sum(b.num_rows for b in batches)
    out_a, out_p = np.empty(max(total, 1), np.int32), np.empty(max(total, 1), np.int32)
    part = (C.c_int64 * max(n_partitions, 1))()
    fn = lib().orc_q2_collect
    times = []
    m = 0
    for _ in range(max(repeat, 1)):
        t = time.perf_counter()
        m = fn(au, pr, rows, C.c_int32(n), C.c_int64(modulus), C.c_int64(rhs), C.c_int32(n_partitions), C.c_int32(n_threads),
               out_a.ctypes.data_as(C.c_void_p), out_p.ctypes.data_as(C.c_void_p), part)
        times.append(time.perf_counter() - t)
    tbl = pa.table({"auction": pa.array(out_a[:m]), "price": pa.array(out_p[:m])})
    return tbl, times


def q8_collect(persons: list[pa.RecordBatch], auctions: list[pa.RecordBatch], n_partitions: int, n_threads: int, repeat: int = 1):
    """NEXMark q8 over one window by the native partition-parallel pipeline (orc_q8_collect: Partial DISTINCT per input
    partition -> hash repartition -> FinalPartitioned DISTINCT -> partitioned join; native threads, no Python per batch).
    Returns (result table (p_id, name), seconds per call as a list of `repeat` wall times)."""
    import time
    fn = lib().orc_q8_collect
    fn.restype = C.c_int64
    persons = [b for b in persons if b.num_rows]
    auctions = [b for b in auctions if b.num_rows]
    npb, nab = len(persons), len(auctions)
    pid = (C.c_void_p * max(npb, 1))(*[b.column("p_id").buffers()[1].address + 4 * b.column("p_id").offset for b in persons])
    noff = (C.c_void_p * max(npb, 1))(*[b.column("name").buffers()[1].address + 4 * b.column("name").offset for b in persons])
    ndat = (C.c_void_p * max(npb, 1))(*[(b.column("name").buffers()[2].address if b.column("name").buffers()[2] is not None else 0) for b in persons])
    prow = (C.c_int64 * max(npb, 1))(*[b.num_rows for b in persons])
    sel = (C.c_void_p * max(nab, 1))(*[b.column("seller").buffers()[1].address + 4 * b.column("seller").offset for b in auctions])
    arow = (C.c_int64 * max(nab, 1))(*[b.num_rows for b in auctions])
    rows = sum(b.num_rows for b in persons)
    nbytes = sum(b.column("name").buffers()[2].size if b.column("name").buffers()[2] is not None else 0 for b in persons)
    out_pid, out_off, out_dat = np.empty(max(rows, 1), np.int32), np.empty(rows + 1, np.int32), np.empty(max(nbytes, 1), np.uint8)
    times, m = [], 0
    for _ in range(max(repeat, 1)):
        t = time.perf_counter()
        m = fn(pid, noff, ndat, prow, C.c_int32(npb), sel, arow, C.c_int32(nab), C.c_int32(n_partitions), C.c_int32(n_threads),
               out_pid.ctypes.data_as(C.c_void_p), out_off.ctypes.data_as(C.c_void_p), out_dat.ctypes.data_as(C.c_void_p))
        times.append(time.perf_counter() - t)
    names = pa.Array.from_buffers(pa.utf8(), m, [None, pa.py_buffer(out_off[:m + 1].tobytes()), pa.py_buffer(out_dat[:int(out_off[m]) if m else 0].tobytes())])
    return pa.table({"p_id": pa.array(out_pid[:m]), "name": names}), times


def q5_collect(bids: list[pa.RecordBatch], n_partitions: int, n_threads: int, repeat: int = 1):
    """NEXMark q5 over one window by the native partition-parallel pipeline (orc_q5_collect).  Returns (table (auction, num),
    seconds per call)."""
    import time
    fn = lib().orc_q5_collect
    fn.restype = C.c_int64
    bids = [b for b in bids if b.num_rows]
    n = len(bids)
    au = (C.c_void_p * max(n, 1))(*[b.column("auction").buffers()[1].address + 4 * b.column("auction").offset for b in bids])
    rows = (C.c_int64 * max(n, 1))(*[b.num_rows for b in bids])
    cap = 1 << 16
    out_a, out_n = np.empty(cap, np.int32), np.empty(cap, np.uint64)
    times, m = [], 0
    for _ in range(max(repeat, 1)):
        t = time.perf_counter()
        m = fn(au, rows, C.c_int32(n), C.c_int32(n_partitions), C.c_int32(n_threads), out_a.ctypes.data_as(C.c_void_p), out_n.ctypes.data_as(C.c_void_p),
               C.c_int64(cap))
        times.append(time.perf_counter() - t)
    return pa.table({"auction": pa.array(out_a[:m]), "num": pa.array(out_n[:m])}), times


def _col_ptrs(batches, name, width=4):
    return (C.c_void_p * max(len(batches), 1))(*[b.column(name).buffers()[1].address + width * b.column(name).offset for b in batches])


def _utf8_ptrs(batches, name):
    offs = (C.c_void_p * max(len(batches), 1))(*[b.column(name).buffers()[1].address + 4 * b.column(name).offset for b in batches])
    data = (C.c_void_p * max(len(batches), 1))(*[(b.column(name).buffers()[2].address if b.column(name).buffers()[2] is not None else 0) for b in batches])
    return offs, data


def q1_collect(bids: list[pa.RecordBatch], n_partitions: int, n_threads: int, repeat: int = 1):
    """NEXMark q1 natively (orc_q1_collect): the computed column 0.908 * CAST(price AS Float64) batch by batch, the pass-through
    columns shared like the reference's Arc clones.  Returns (list of result batches, seconds per call)."""
    import time
    fn = lib().orc_q1_collect
    fn.restype = None
    n = len(bids)
    price = _col_ptrs(bids, "price")
    rows = (C.c_int64 * max(n, 1))(*[b.num_rows for b in bids])
    outs = [np.empty(b.num_rows, np.float64) for b in bids]
    optr = (C.c_void_p * max(n, 1))(*[o.ctypes.data for o in outs])
    times = []
    for _ in range(max(repeat, 1)):
        t = time.perf_counter()
        fn(price, rows, C.c_int32(n), C.c_int32(n_partitions), C.c_int32(n_threads), optr)
        times.append(time.perf_counter() - t)
    res = [pa.RecordBatch.from_arrays([b.column("auction"), b.column("bidder"), pa.array(o), b.column("b_date_time")], names=["auction", "bidder", "price", "b_date_time"])
           for b, o in zip(bids, outs)]
    return res, times


def q3_collect(auctions: list[pa.RecordBatch], persons: list[pa.RecordBatch], n_partitions: int, n_threads: int, repeat: int = 1):
    """NEXMark q3 natively (orc_q3_collect: the two filters -> hash repartition on the join key -> partitioned hash join ->
    projection).  Returns (table (name, city, state, a_id), seconds per call)."""
    import time
    fn = lib().orc_q3_collect
    fn.restype = C.c_int64
    auctions = [b for b in auctions if b.num_rows]
    persons = [b for b in persons if b.num_rows]
    na, npb = len(auctions), len(persons)
    a_id, a_sel, a_cat = _col_ptrs(auctions, "a_id"), _col_ptrs(auctions, "seller"), _col_ptrs(auctions, "category")
    a_rows = (C.c_int64 * max(na, 1))(*[b.num_rows for b in auctions])
    p_id = _col_ptrs(persons, "p_id")
    (n_o, n_d), (c_o, c_d), (s_o, s_d) = _utf8_ptrs(persons, "name"), _utf8_ptrs(persons, "city"), _utf8_ptrs(persons, "state")
    p_rows = (C.c_int64 * max(npb, 1))(*[b.num_rows for b in persons])
    cap = max(1024, 2 * sum(b.num_rows for b in auctions))
    while True:
        bcap = cap * 48
        o_a = np.empty(cap, np.int32)
        o_no, o_co, o_so = (np.empty(cap + 1, np.int32) for _ in range(3))
        o_n, o_c, o_s = (np.empty(bcap, np.uint8) for _ in range(3))
        times, m = [], 0
        for _ in range(max(repeat, 1)):
            t = time.perf_counter()
            m = fn(a_id, a_sel, a_cat, a_rows, C.c_int32(na), p_id, n_o, n_d, c_o, c_d, s_o, s_d, p_rows, C.c_int32(npb), C.c_int32(n_partitions), C.c_int32(n_threads),
                   o_a.ctypes.data_as(C.c_void_p), o_no.ctypes.data_as(C.c_void_p), o_n.ctypes.data_as(C.c_void_p), o_co.ctypes.data_as(C.c_void_p),
                   o_c.ctypes.data_as(C.c_void_p), o_so.ctypes.data_as(C.c_void_p), o_s.ctypes.data_as(C.c_void_p), C.c_int64(cap), C.c_int64(bcap))
            times.append(time.perf_counter() - t)
            if m < 0:
                break
        if m >= 0:
            break
        cap *= 4
    utf8 = lambda off, dat: pa.Array.from_buffers(pa.utf8(), m, [None, pa.py_buffer(off[:m + 1].tobytes()), pa.py_buffer(dat[:int(off[m]) if m else 0].tobytes())])
    return pa.table({"name": utf8(o_no, o_n), "city": utf8(o_co, o_c), "state": utf8(o_so, o_s), "a_id": pa.array(o_a[:m])}), times


def _schema_from_json(s: dict) -> pa.Schema:
    def ty(t):
        if isinstance(t, dict):
            unit, tz = t["Timestamp"]
            return pa.timestamp({"Second": "s", "Millisecond": "ms", "Microsecond": "us", "Nanosecond": "ns"}[unit], tz)
        return {"Int32": pa.int32(), "UInt32": pa.uint32(), "Int64": pa.int64(), "UInt64": pa.uint64(), "Float64": pa.float64(), "Utf8": pa.utf8()}[t]
    return pa.schema([pa.field(f["name"], ty(f["data_type"]), f.get("nullable", False)) for f in s["fields"]], metadata=s.get("metadata") or None)


def _compare_schema(a: list[str], b: list[str]) -> bool:
    sup, sub = (a, b) if len(a) >= len(b) else (b, a)
    return all(x in set(sup) for x in sub)


class PlanExecutor:
    """flock::runtime::context::ExecutionContext over the oracle operators."""

    def __init__(self, plans, threads: int = 1):
        plans = json.loads(plans) if isinstance(plans, str) else plans
        if isinstance(plans, dict) and "plan" in plans:
            plans = plans["plan"]["execution_plans"]
        self.plans = plans if isinstance(plans, list) else [plans]
        self.fed: dict[int, Partitions] = {}
        self.pool = ThreadPoolExecutor(threads) if threads > 1 else None

    # ---- feeding (context.rs:257-325)
    def _leaves(self):
        queue, out = list(self.plans), []
        while queue:
            p = queue.pop(0)
            ch = [p[k] for k in ("input", "left", "right") if k in p]
            if not ch:
                out.append(p)
            queue += ch
        return out

    @staticmethod
    def _projected_names(leaf: dict) -> list[str]:
        names = [f["name"] for f in leaf["schema"]["fields"]]
        proj = leaf.get("projection")
        if proj is not None and all(0 <= p < len(names) for p in proj):
            return [names[p] for p in proj]
        return names

    def feed_data_sources(self, sources: list[Partitions]) -> None:
        sources = list(sources)
        self.fed = {}
        for leaf in self._leaves():
            want = self._projected_names(leaf)
            for i, src in enumerate(sources):
                first = next((b for part in src for b in part), None)
                if first is not None and _compare_schema(want, first.schema.names):
                    self.fed[id(leaf)] = sources.pop(i)
                    break

    def clean_data_sources(self) -> None:
        self.fed = {}

    # ---- execution
    def _map(self, fn, items):
        if self.pool is not None and len(items) > 1:
            return list(self.pool.map(fn, items))
        return [fn(x) for x in items]

    def _exec(self, p: dict) -> Partitions:
        tag = p["execution_plan"]
        if tag == "memory_exec":
            want = self._projected_names(p)
            parts = self.fed.get(id(p))
            if parts is None:
                sch = _schema_from_json(p["schema"])
                sch = pa.schema([sch.field(n) for n in want], metadata=sch.metadata)
                return [[pa.RecordBatch.from_arrays([pa.array([], f.type) for f in sch], schema=sch)]]
            return [[b.select(want) for b in part] for part in parts]
        if tag == "repartition_exec":
            parts = self._exec(p["input"])
            part = p["partitioning"]
            if "RoundRobinBatch" in part:
                n = part["RoundRobinBatch"]
                out = [[] for _ in range(n)]
                k = 0
                for src in parts:
                    for b in src:
                        out[k % n].append(b)
                        k += 1
                return out
            exprs, n = part["Hash"]
            out = [[] for _ in range(n)]

            def split(b):
                keys = [_resolve(b.schema.names, e) for e in exprs]
                pid = partition_ids(b, keys, n)
                return [b.take(pa.array(np.nonzero(pid == q)[0])) for q in range(n)]
            for src in parts:
                for pieces in self._map(split, [b for b in src if b.num_rows]):
                    for q, piece in enumerate(pieces):
                        if piece.num_rows:
                            out[q].append(piece)
            return out
        if tag == "coalesce_batches_exec":
            target = p.get("target_batch_size", 4096)
            out = []
            for src in self._exec(p["input"]):
                res, buf, rows = [], [], 0
                for b in src:
                    if b.num_rows == 0:
                        continue
                    buf.append(b)
                    rows += b.num_rows
                    if rows >= target:
                        res.append(_concat(buf, b.schema))
                        buf, rows = [], 0
                if buf:
                    res.append(_concat(buf, buf[0].schema))
                out.append(res if res else src[:1])
            return out
        if tag in ("coalesce_partitions_exec", "merge_exec"):
            return [[b for src in self._exec(p["input"]) for b in src]]
        if tag == "filter_exec":
            return [self._map(lambda b: filter_batch(b, p["predicate"]), src) for src in self._exec(p["input"])]
        if tag == "projection_exec":
            exprs = [(e, n) for e, n in p["expr"]]
            return [self._map(lambda b: project_batch(b, exprs), src) for src in self._exec(p["input"])]
        if tag == "hash_aggregate_exec":
            parts = self._exec(p["input"])
            mode = p["mode"]

            def run(src):
                schema = src[0].schema if src else None
                if schema is None:
                    return []
                batch = _concat([b for b in src if b.num_rows], schema)
                names = batch.schema.names
                group = [(_resolve(names, e), n) for e, n in p["group_expr"]]
                aggrs, state = [], len(group)
                for a in p["aggr_expr"]:
                    f = a["aggregate_expr"]
                    if mode in ("Final", "FinalPartitioned"):
                        aggrs.append({"func": f, "col": state, "name": a["name"]})
                        state += len(_AGG_STATE[f])
                    else:
                        e = a.get("expr")
                        col = _resolve(names, e) if e and e["physical_expr"] == "column" else -1
                        aggrs.append({"func": f, "col": col, "name": a["name"]})
                return [hash_aggregate(batch, mode, group, aggrs)]
            return self._map(run, parts)
        if tag == "hash_join_exec":
            lparts, rparts = self._exec(p["left"]), self._exec(p["right"])
            if p.get("mode", "Partitioned") == "CollectLeft" or len(lparts) != len(rparts):
                lparts = [[b for src in lparts for b in src]] * len(rparts)

            def on_idx(o, names):
                return names.index(o) if isinstance(o, str) else _resolve(names, o)

            def run(pair):
                lsrc, rsrc = pair
                if not lsrc or not rsrc:
                    return []
                build = _concat([b for b in lsrc if b.num_rows], lsrc[0].schema)
                lk = [on_idx(l, build.schema.names) for l, _ in p["on"]]
                out = []
                for rb in rsrc:
                    rk = [on_idx(r, rb.schema.names) for _, r in p["on"]]
                    out.append(hash_join(build, rb, lk, rk))
                return out
            return self._map(run, list(zip(lparts, rparts)))
        if tag == "sort_exec":
            # SortExec (DataFusion 6): sorts each input partition by the expr list (planners put a CoalescePartitions /
            # MergeExec below it).  Rows with equal sort keys have NO defined order in the reference (arrow
            # lexsort_to_indices is not stable); the oracle breaks ties by the remaining columns, ascending, so that
            # two independent implementations can agree -- one valid instance of the reference's behaviour.
            return [[sort_batch(_concat([b for b in src if b.num_rows], src[0].schema), p["expr"])] if src else [] for src in self._exec(p["input"])]
        if tag == "global_limit_exec":
            src = [b for part in self._exec(p["input"]) for b in part]
            if not src:
                return [[]]
            whole = _concat(src, src[0].schema)
            return [[whole.slice(0, p["limit"])]]
        if tag == "window_agg_exec":
            # WindowAggExec with ROW_NUMBER() OVER (PARTITION BY .. ORDER BY ..): the input arrives sorted by
            # (partition keys, order keys) -- the planner's SortExec below -- and the window columns are emitted FIRST,
            # then the input columns (q6_plan.fmt).  Row order is the input order.
            return [[window_batch(_concat([b for b in src if b.num_rows], src[0].schema), p["window_expr"])] if src else []
                    for src in self._exec(p["input"])]
        raise OracleError(f"oracle: execution plan node {tag} is not restated")

    def execute_partitioned(self) -> list[Partitions]:
        return [self._exec(p) for p in self.plans]

    def execute(self) -> list[list[pa.RecordBatch]]:
        """collect(): all partitions of every plan, merged (context.rs:172-191)."""
        return [[b for part in self._exec(p) for b in part] for p in self.plans]


def execute_plan(plan, sources: list[Partitions], threads: int = 1) -> pa.Table:
    """Feeds `sources` to `plan` and returns the collected result of plan 0 as one table."""
    ex = PlanExecutor(plan, threads)
    ex.feed_data_sources(sources)
    batches = [b for b in ex.execute()[0]]
    nonempty = [b for b in batches if b.num_rows]
    if nonempty:
        return pa.Table.from_batches(nonempty)
    return pa.Table.from_batches(batches[:1]) if batches else pa.table({})


def canonical(table: pa.Table) -> pa.Table:
    """Rows in a canonical order: the comparator of assert_batches_sorted_eq! (flock/src/test_util.rs:60-90)."""
    t = table.combine_chunks()
    if t.num_rows == 0 or t.num_columns == 0:
        return t
    keys = [(f"c{i}", "ascending") for i in range(t.num_columns)]
    tmp = pa.table({f"c{i}": t.column(i) for i in range(t.num_columns)})
    order = pa.compute.sort_indices(tmp, sort_keys=keys)
    return t.take(order)


def assert_tables_equal(actual: pa.Table, expected: pa.Table, sort: bool = True, check_names: bool = True) -> None:
    """Bit-exact comparison of schema and rows (after the canonical sort unless the plan orders its output)."""
    if check_names:
        assert actual.schema.names == expected.schema.names, f"column names differ: {actual.schema.names} vs {expected.schema.names}"
    assert [f.type for f in actual.schema] == [f.type for f in expected.schema], \
        f"column types differ: {[str(f.type) for f in actual.schema]} vs {[str(f.type) for f in expected.schema]}"
    assert actual.num_rows == expected.num_rows, f"row counts differ: {actual.num_rows} vs {expected.num_rows}"
    a, e = (canonical(actual), canonical(expected)) if sort else (actual.combine_chunks(), expected.combine_chunks())
    for i in range(a.num_columns):
        ca, ce = a.column(i), e.column(i)
        if pa.types.is_floating(ca.type):
            va = ca.to_numpy(zero_copy_only=False).view(np.int64) if ca.null_count == 0 else None
            ve = ce.to_numpy(zero_copy_only=False).view(np.int64) if ce.null_count == 0 else None
            if va is not None and ve is not None:
                assert np.array_equal(va, ve), f"column {a.schema.names[i]}: float bit patterns differ"
                continue
        assert ca.equals(ce), f"column {a.schema.names[i]} differs"
