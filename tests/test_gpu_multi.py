"""Multi-GPU parity (`-m gpu`, needs >= 2 devices; skipped on a 1-GPU box): two processes, one GPU each,
an NCCL communicator attached through the C ABI (flockgpu_comm_init).  Checks the all-to-all itself and the
distributed execution of the reference's plans (every RepartitionExec(Hash) becomes an NVLink all-to-all):
the union of the ranks' results must equal the CPU oracle's single-process result."""
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


def _worker(rank, world, port, out_dir):
    import torch.distributed as dist
    import flock_b200 as fb
    from flock_b200 import nexgen, plans, sharding
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
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
    got = ctx.hash_exchange(t, [0]).to_arrow()
    pid = sharding.partition_ids(b, [0], world)
    mine = b.filter(pa.array(pid == rank))
    assert got.filter(pa.compute.equal(got["src"], rank)).to_batches()[0].equals(mine) if mine.num_rows else True   # own rows, input order kept
    assert np.all(sharding.partition_ids(got.combine_chunks().to_batches()[0], [0], world) == rank)                # only keys routed to me
    counts = [None] * world
    dist.all_gather_object(counts, (b.num_rows, got.num_rows))
    assert sum(c[0] for c in counts) == sum(c[1] for c in counts)                                                   # nothing lost

    # ---- distributed plans
    ev = nexgen.generate(400_000, seed=21, batch_rows=4096)
    for q in ("q8", "q5", "q3"):
        ec = fb.ExecutionContext(ctx, plans.QUERIES[q]())
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
def test_two_gpu_exchange_and_plans(tmp_path):
    import torch.multiprocessing as mp
    import oracle
    from flock_b200 import nexgen, plans
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    ev = nexgen.generate(400_000, seed=21, batch_rows=4096)
    for q in ("q8", "q5", "q3"):
        parts = [pa.ipc.open_stream((tmp_path / f"{q}_rank{r}.arrow").read_bytes()).read_all() for r in range(world)]
        got = pa.concat_tables(parts)
        want = oracle.execute_plan(plans.QUERIES[q](), [[ev[r]] for r in plans.SOURCES[q]])
        oracle.assert_tables_equal(got, want)
