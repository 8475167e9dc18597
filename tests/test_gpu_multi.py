"""Multi-GPU parity (`-m gpu`, needs >= 2 devices; skipped on a 1-GPU box): two processes, one GPU each, a communicator
attached through the C ABI (flockgpu_comm_init).  Checks the exchange itself -- once over NVLink peer windows (the
product path: the partition kernel stores into the receiver's HBM) and once over the NCCL fallback -- and the
distributed execution of the reference's plans (every RepartitionExec(Hash) becomes an exchange): the union of the
ranks' results must equal the CPU oracle's single-process result.  bench.py --gpus N repeats the q8 check on the
driver's multi-GPU box (its JSON line carries `parity_check`)."""
import io
import os
import socket

import numpy as np
import pyarrow as pa
import pytest

pytestmark = pytest.mark.gpu


def _n_gpus() -> int:
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


def _ipc(t: pa.Table) -> bytes:
    sink = io.BytesIO()
    with pa.ipc.new_stream(sink, t.schema) as w:
        w.write_table(t)
    return sink.getvalue()


def _worker(rank, world, port, out_dir, mode):
    """Every rank leaves its own traceback behind: mp.spawn reports one failing process only, and usually the one
    that merely lost its peer."""
    try:
        _worker_body(rank, world, port, out_dir, mode)
    except BaseException:
        import traceback
        with open(os.path.join(out_dir, f"rank{rank}.err"), "w") as f:
            traceback.print_exc(file=f)
        raise


def _worker_body(rank, world, port, out_dir, mode):
    import torch.distributed as dist
    import flock_b200 as fb
    from flock_b200 import nexgen, plans, sharding
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    if mode == "nccl":
        os.environ["FLOCKGPU_EXCHANGE"] = "nccl"        # read by flockgpu_comm_init: peer windows off
    os.environ["FLOCKGPU_WINDOW_MB"] = "256"
    dist.init_process_group("gloo", rank=rank, world_size=world)          # control plane only: ships the NCCL id
    ctx = fb.Context(rank)
    ids = [fb.Context.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    ctx.comm_init(ids[0], rank, world)

    # ---- the exchange itself: mixed fixed-width + Utf8 columns, ragged partition sizes
    rng = np.random.default_rng(100 + rank)
    n = 50_000 + 1000 * rank
    words = ["", "a", "portland", "san francisco", "x" * 40]
    b = pa.RecordBatch.from_arrays([pa.array(rng.integers(0, 1 << 30, n).astype(np.int32)), pa.array([words[k] for k in rng.integers(0, 5, n)]),
                                    pa.array(np.full(n, rank, np.int64))], names=["k", "s", "src"])
    t = ctx.import_batches([b])
    dist.barrier()                                   # a rank that is late by more than the exchange time-out fails the others
    held = ctx.hash_exchange(t, [0])                 # stays alive: its columns are views into the receive window
    got = held.to_arrow()
    pid = sharding.partition_ids(b, [0], world)
    mine = b.filter(pa.array(pid == rank))
    assert got.filter(pa.compute.equal(got["src"], rank)).to_batches()[0].equals(mine) if mine.num_rows else True   # own rows, input order kept
    assert np.all(sharding.partition_ids(got.combine_chunks().to_batches()[0], [0], world) == rank)                # only keys routed to me
    counts = [None] * world
    dist.all_gather_object(counts, (b.num_rows, got.num_rows))
    assert sum(c[0] for c in counts) == sum(c[1] for c in counts)                                                   # nothing lost
    # a second exchange while the first result is still alive lands behind it in the window and leaves it intact;
    # routing on (k, s) uses the fixed-width column alone, so the result is the same relation
    again = ctx.hash_exchange(t, [0, 1]).to_arrow()
    assert again.equals(got) and held.to_arrow().equals(got)
    # the operators run on a received relation like on any other
    agg = ctx.hash_aggregate(held, [2], [("count", -1, "n")], "single").to_arrow()
    per_src = {s: got.filter(pa.compute.equal(got["src"], s)).num_rows for s in range(world)}
    assert dict(zip(agg["src"].to_pylist(), agg["n"].to_pylist())) == {s: n for s, n in per_src.items() if n}
    del held

    # ---- NULLs cross the exchange with their rows: rank 0 holds NULLs (key and payload), rank 1 the same nullable
    # schema without a single NULL (every rank lays out the same buffers); rows with a NULL key meet on one rank
    nn = 30_000 + 7 * rank
    holes = rank == 0 or mode == "nccl"              # (the fallback refuses per rank: give every rank a NULL so none waits for a peer that raised)
    mask_k = (np.arange(nn) % 9 == 0) if holes else np.zeros(nn, bool)
    mask_s = (np.arange(nn) % 4 == 1) if holes else np.zeros(nn, bool)
    nb = pa.RecordBatch.from_arrays([pa.array(rng.integers(0, 5000, nn).astype(np.int32), mask=mask_k),
                                     pa.array([None if m else words[k] for k, m in zip(rng.integers(0, 5, nn), mask_s)], pa.utf8()),
                                     pa.array(rng.integers(0, 1 << 40, nn), mask=mask_s[::-1].copy()), pa.array(np.full(nn, rank, np.int32))],
                                    names=["k", "s", "v", "src"])
    dist.barrier()
    if mode == "peer":
        moved = ctx.hash_exchange(ctx.import_batches([nb]), [0]).to_arrow()
        everything = [None] * world
        dist.all_gather_object(everything, (_ipc(pa.Table.from_batches([nb])), _ipc(moved)))
        if rank == 0:
            sent = pa.concat_tables([pa.ipc.open_stream(e[0]).read_all() for e in everything])
            landed = [pa.ipc.open_stream(e[1]).read_all() for e in everything]
            import oracle
            oracle.assert_tables_equal(pa.concat_tables(landed), sent)
            homes = {}
            for r, part in enumerate(landed):
                for k in set(part["k"].to_pylist()):
                    assert homes.setdefault(k, r) == r
            assert None in homes
    else:
        with pytest.raises(fb.FlockGpuError) as info:       # the NCCL fallback moves values only and says so
            ctx.hash_exchange(ctx.import_batches([nb]), [0])
        assert info.value.code == -2

    # ---- a global aggregate with one EMPTY shard: rank 1 scans no bids; MAX must come out of rank 0's state alone, and
    # with every shard empty the merged MAX is NULL while COUNT is 0
    bids = nexgen.split_batches(nexgen.bids(50_000, seed=9), 8192)
    mx = plans.aggregate_expr("max", "MAX(bid.price)", plans.column("price", 2), "Int32")
    cnt = plans.aggregate_expr("count", "COUNT(UInt8(1))", plans.literal("UInt8", 1), "UInt64")
    plan = plans.two_phase_aggregate([], [mx, cnt], plans.repartition_rr(plans.memory_exec(nexgen.bid_schema(), [0, 1, 2, 3])))
    for shard, want in ((bids if rank == 0 else [bids[0].slice(0, 0)], "value"), ([bids[0].slice(0, 0)], "null")):
        ec = fb.ExecutionContext(ctx, plan)
        dist.barrier()
        ec.feed_data_sources([[shard]])
        out = pa.Table.from_batches(ec.execute()[0])
        ec.close()
        if rank == 0:
            price = np.concatenate([x["price"].to_numpy() for x in bids])
            if want == "value":
                assert out.num_rows == 1 and out.column(0).to_pylist() == [int(price.max())] and out.column(1).to_pylist() == [price.size]
            else:
                assert out.num_rows == 1 and out.column(0).to_pylist() == [None] and out.column(1).to_pylist() == [0]
        else:
            assert out.num_rows == 0

    # ---- distributed plans
    ev = nexgen.generate(400_000, seed=21, batch_rows=4096)
    for q in ("q8", "q5", "q3"):
        ec = fb.ExecutionContext(ctx, plans.QUERIES[q]())
        dist.barrier()
        ec.feed_data_sources([[sharding.round_robin(ev[r], rank, world)] for r in plans.SOURCES[q]])
        out = pa.Table.from_batches(ec.execute()[0])
        ec.close()
        with open(os.path.join(out_dir, f"{q}_rank{rank}.arrow"), "wb") as f:
            f.write(_ipc(out))
    dist.barrier()
    ctx.close()
    dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.skipif(_n_gpus() < 2, reason="needs two GPUs")
@pytest.mark.timeout(900)
@pytest.mark.parametrize("mode", ["peer", "nccl"])
def test_two_gpu_exchange_and_plans(tmp_path, mode):
    import torch.multiprocessing as mp
    import oracle
    from flock_b200 import nexgen, plans
    world = 2
    try:
        mp.spawn(_worker, args=(world, _free_port(), str(tmp_path), mode), nprocs=world, join=True)
    except Exception as e:
        traces = "\n".join(f"---- {f.name}\n{f.read_text()}" for f in sorted(tmp_path.glob("rank*.err")))
        raise AssertionError(f"{e}\n{traces}") from None
    ev = nexgen.generate(400_000, seed=21, batch_rows=4096)
    for q in ("q8", "q5", "q3"):
        parts = [pa.ipc.open_stream((tmp_path / f"{q}_rank{r}.arrow").read_bytes()).read_all() for r in range(world)]
        got = pa.concat_tables(parts)
        want = oracle.execute_plan(plans.QUERIES[q](), [[ev[r]] for r in plans.SOURCES[q]])
        oracle.assert_tables_equal(got, want)
