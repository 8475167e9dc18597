"""The N > 1 path on CPU ranks (`-m "not gpu"`): world_size-2 `gloo` run of the distributed q8 shape
(SURVEY.md section 8e): round-robin input sharding -> local Partial (distinct) -> hash exchange routed by
flock_b200.sharding.partition_ids (the numpy restatement of the DEVICE routing function of partition.cu)
-> FinalPartitioned -> partition-wise join.  The union of the ranks' results must equal the
single-process oracle result, and every key must live on exactly one rank."""
import io
import os
import socket

import numpy as np
import pyarrow as pa
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle
from flock_b200 import nexgen, plans, sharding

N_EVENTS, SEED = 200_000, 11


def _ipc(batch: pa.RecordBatch) -> bytes:
    sink = io.BytesIO()
    with pa.ipc.new_stream(sink, batch.schema) as w:
        w.write_batch(batch)
    return sink.getvalue()


def _from_ipc(buf: bytes) -> pa.RecordBatch:
    t = pa.ipc.open_stream(buf).read_all().combine_chunks()
    return t.to_batches()[0] if t.num_rows else pa.RecordBatch.from_arrays([pa.array([], f.type) for f in t.schema], schema=t.schema)


def _exchange(batch: pa.RecordBatch, key_cols, rank, world):
    """hash_partition(world) + all-to-all: returns the rows every rank routed to `rank`."""
    pid = sharding.partition_ids(batch, key_cols, world)
    pieces = [_ipc(batch.filter(pa.array(pid == r))) for r in range(world)]
    gathered = [None] * world
    dist.all_gather_object(gathered, pieces)          # gathered[s][r] = what rank s sends to rank r
    mine = [_from_ipc(gathered[s][rank]) for s in range(world)]
    return pa.Table.from_batches(mine).combine_chunks().to_batches()[0]


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ev = nexgen.generate(N_EVENTS, seed=SEED, batch_rows=2048, relations=("person", "auction"))
    persons = sharding.round_robin(ev["person"], rank, world)
    auctions = sharding.round_robin(ev["auction"], rank, world)
    p_local = pa.Table.from_batches(persons).combine_chunks().to_batches()[0].select(["p_id", "name"])
    a_local = pa.Table.from_batches(auctions).combine_chunks().to_batches()[0].select(["seller"])
    # Partial (distinct) before the exchange, so that groups travel, not rows
    p_part = oracle.hash_aggregate(p_local, "Partial", [(0, "p_id"), (1, "name")], [])
    a_part = oracle.hash_aggregate(a_local, "Partial", [(0, "seller")], [])
    p_recv = _exchange(p_part, [0], rank, world)      # routed by p_id
    a_recv = _exchange(a_part, [0], rank, world)      # routed by seller: equal keys of both sides meet on one rank
    P = oracle.hash_aggregate(p_recv, "FinalPartitioned", [(0, "p_id"), (1, "name")], [])
    A = oracle.hash_aggregate(a_recv, "FinalPartitioned", [(0, "seller")], [])
    joined = oracle.hash_join(P, A, [0], [0]).select(["p_id", "name"])
    with open(os.path.join(out_dir, f"rank{rank}.arrow"), "wb") as f:
        f.write(_ipc(joined))
    # every p_id this rank owns routes to this rank
    assert np.all(sharding.partition_ids(P, [0], world) == rank)
    dist.barrier()
    dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.timeout(300)
def test_q8_exchange_on_two_gloo_ranks(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    parts = [_from_ipc((tmp_path / f"rank{r}.arrow").read_bytes()) for r in range(world)]
    got = pa.Table.from_batches(parts)
    ev = nexgen.generate(N_EVENTS, seed=SEED, batch_rows=2048, relations=("person", "auction"))
    want = oracle.execute_plan(plans.q8(), [[ev["person"]], [ev["auction"]]])
    oracle.assert_tables_equal(got, want)
    assert all(p.num_rows > 0 for p in parts)                      # both ranks own part of the result
    ids = [set(p["p_id"].to_pylist()) for p in parts]
    assert not (ids[0] & ids[1])                                     # a key lives on exactly one rank


def test_routing_function_properties():
    rng = np.random.default_rng(0)
    b = pa.RecordBatch.from_arrays([pa.array(rng.integers(1000, 200_000, 100_000).astype(np.int32)),
                                    pa.array(rng.integers(0, 1 << 62, 100_000))], names=["k32", "k64"])
    for world in (2, 4, 8):
        pid = sharding.partition_ids(b, [0], world)
        assert pid.min() >= 0 and pid.max() == world - 1
        counts = np.bincount(pid, minlength=world)
        assert counts.min() > 0.8 * counts.mean()                   # the radix split is balanced
        k = b["k32"].to_numpy()
        first = {}
        for key, p in zip(k[:5000], pid[:5000]):
            assert first.setdefault(int(key), int(p)) == int(p)    # deterministic per key
    # an Int32 key and the same values as the hi:lo pair pack differently but consistently
    assert sharding.pack_keys(b, [0]).dtype == np.uint64
    two = pa.RecordBatch.from_arrays([b["k32"], b["k32"]], names=["a", "b"])
    packed = sharding.pack_keys(two, [0, 1])
    assert np.array_equal(packed >> np.uint64(32), packed & np.uint64(0xFFFFFFFF))
    assert sharding.round_robin(list(range(10)), 1, 4) == [1, 5, 9]
