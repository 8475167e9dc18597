#!/usr/bin/env python
"""tools/nexmark_bench.py -- the five BASELINE.json configurations through the GPU executor, one JSON line each.

    python tools/nexmark_bench.py [--queries q1,q2,q3,q5,q8] [--scale 1.0] [--reps 20] [--check]
    python -m torch.distributed.run --nproc-per-node N ... tools/nexmark_bench.py --queries q8 --scale 1.0

Per query it reports events/s with the relations resident in HBM (CUDA events around whole plan executions,
median and best of --reps), the end-to-end figure with host batches, the per-kernel CUDA-event profile, the
algorithmic bytes of SURVEY.md section 8(d) and the resulting fraction of the measured HBM peak, plus the
CPU oracle on the host cores on the same input (skipped with --no-cpu or when the input exceeds --cpu-max-rows).
With WORLD_SIZE > 1 (torchrun) every rank scans its round-robin share of the batches and the plan's Hash
repartitions become NCCL all-to-alls; rank 0 prints max-over-ranks times.  bench.py stays the contract bench
(q2); this tool is the survey table behind profiles/.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import numpy as np      # noqa: E402
import pyarrow as pa    # noqa: E402

import flock_b200 as fb                                   # noqa: E402
from flock_b200 import nexgen, plans, sharding            # noqa: E402

BATCH = 65536
# BASELINE.json configs: events of the stream per query (scale 1.0)
EVENTS = {"q1": None, "q2": None, "q3": 10_000_000, "q5": None, "q8": 1_000_000_000}
BIDS = {"q1": 65536, "q2": 10_000_000, "q5": 100_000_000}


# Libraries write to fd 1 behind Python's back (NCCL prints "NCCL version ..." there): keep the real stdout for the
# JSON line(s) only and point fd 1 at stderr for everything else.
_JSON_OUT = os.fdopen(os.dup(1), "w")
os.dup2(2, 1)


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def peak_gbs():
    p = ROOT / "MEASURED_PEAKS.json"
    return float(json.loads(p.read_text())["hbm_gbs"]) if p.exists() else 6650.0


def make_inputs(q: str, scale: float, seed: int, rank: int, world: int) -> dict:
    """relation name -> list of 64 Ki-row batches this rank scans."""
    cols = {"person": ["p_id", "name", "city", "state"] if q == "q3" else ["p_id", "name"],
            "auction": ["a_id", "seller", "category"] if q == "q3" else ["seller"]}
    if world > 1 and not os.environ.get("NEXMARK_ROUND_ROBIN"):
        # every rank generates ONLY its own contiguous slice of the stream (as if it hosted that slice of the source):
        # a 1 B-event q8 run must not build 20 M persons + 60 M auctions eight times on one host
        def own(total):
            share = (total + world - 1) // world
            first = min(rank * share, total)
            return first, min(share, total - first)

        def pieces(total, fn, extra):
            first, n = own(total)
            parts = [fn(min(4_000_000, n - o), seed, first + o, *extra) for o in range(0, n, 4_000_000)] or [fn(0, seed, first, *extra)]
            tbl = pa.Table.from_batches(parts).combine_chunks()
            return nexgen.split_batches(tbl.to_batches()[0] if tbl.num_rows else parts[0], BATCH)

        if q in BIDS:
            return {"bid": pieces(max(int(BIDS[q] * scale), 1), nexgen.bids, ())}
        n_p, n_a, _ = nexgen.relation_counts(int(EVENTS[q] * scale))
        return {"person": pieces(n_p, nexgen.persons, (cols["person"],)), "auction": pieces(n_a, nexgen.auctions, (cols["auction"],))}
    if q in BIDS:
        n = max(int(BIDS[q] * scale), 1)
        rel = {"bid": nexgen.split_batches(nexgen.bids(n, seed=seed), BATCH)}
    else:
        n_ev = int(EVENTS[q] * scale)
        rel = nexgen.generate(n_ev, seed=seed, batch_rows=BATCH, relations=("person", "auction"), columns=cols)
    return {k: sharding.round_robin(v, rank, world) or [v[0].slice(0, 0)] for k, v in rel.items()}


def algorithmic_bytes(q: str, rel: dict, out_rows: int, extra: dict) -> float:
    """compulsory input columns read + output columns written."""
    rows = {k: sum(b.num_rows for b in v) for k, v in rel.items()}
    def utf8_bytes(batches, name):
        return sum(b[name].nbytes for b in batches) if batches and name in batches[0].schema.names else 0
    if q == "q1":
        return 12.0 * rows["bid"]
    if q == "q2":
        return 4.0 * rows["bid"] + 12.0 * out_rows
    if q == "q3":
        person = 4.0 * rows["person"] + sum(utf8_bytes(rel["person"], c) for c in ("name", "city", "state"))
        return 12.0 * rows["auction"] + person + out_rows * 4.0 + extra.get("out_utf8_bytes", 0)
    if q == "q5":
        return 4.0 * rows["bid"] + 12.0 * extra.get("groups", 0)
    if q == "q8":
        return 4.0 * rows["person"] + utf8_bytes(rel["person"], "name") + 4.0 * rows["auction"] + out_rows * 4.0 + extra.get("out_utf8_bytes", 0)
    return 0.0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--queries", default="q1,q2,q3,q5,q8")
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--q8-scale", type=float, default=None, help="override --scale for q8 (1.0 = 1 B events)")
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--seed", type=int, default=42)
    ap.add_argument("--check", action="store_true", help="compare with the CPU oracle (small scales)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--cpu-max-rows", type=int, default=30_000_000)
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = fb.Context(local)
    if world > 1:
        ids = [fb.Context.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        ctx.comm_init(ids[0], rank, world)

    def dmax(x: float) -> float:
        if dist is None:
            return x
        import torch
        t = torch.tensor([x], dtype=torch.float64, device=f"cuda:{local}")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def dsum(x: float) -> float:
        if dist is None:
            return x
        import torch
        t = torch.tensor([x], dtype=torch.float64, device=f"cuda:{local}")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    peak = peak_gbs()
    for q in args.queries.split(","):
        scale = args.q8_scale if (q == "q8" and args.q8_scale is not None) else args.scale
        t0 = time.time()
        rel = make_inputs(q, scale, args.seed, rank, world)
        order = plans.SOURCES[q]
        resident = {k: ctx.import_batches(v) for k, v in rel.items()}
        rows_in = {k: sum(b.num_rows for b in v) for k, v in rel.items()}
        log(f"[{q} rank {rank}] inputs {rows_in} generated+uploaded in {time.time() - t0:.1f}s")
        ec = fb.ExecutionContext(ctx, plans.QUERIES[q]())

        def run_device():
            ec.feed_tables([resident[r] for r in order])
            return ec.execute_device(0)

        out = run_device()
        out_rows = out.num_rows
        for _ in range(2):
            run_device().num_rows
        times = []
        for _ in range(args.reps):
            ctx.flush_l2()
            if dist is not None:
                dist.barrier()
            ctx.timer_start(0)
            o = run_device()
            ctx.timer_stop(0)
            o.num_rows
            times.append(dmax(ctx.timer_ms(0)))
        ctx.profile_begin()
        run_device().num_rows
        prof = ctx.profile_end()

        e2e_ms = None
        if not args.no_e2e:
            src = [[rel[r]] for r in order]
            ec.feed_data_sources(src); ec.execute(); ec.clean_data_sources()
            ts = []
            for _ in range(max(3, args.reps // 4)):
                t = time.perf_counter()
                ec.feed_data_sources(src)
                res = ec.execute()
                ec.clean_data_sources()
                ts.append(dmax((time.perf_counter() - t) * 1e3))
            e2e_ms = statistics.median(ts)

        res_tbl = out.to_arrow()
        extra = {}
        if q == "q5":
            extra["groups"] = int(np.unique(np.concatenate([b["auction"].to_numpy() for b in rel["bid"]])).size) if rows_in["bid"] <= 120_000_000 else 0
        if q in ("q3", "q8"):
            extra["out_utf8_bytes"] = sum(res_tbl[c].nbytes for c in res_tbl.schema.names if pa.types.is_string(res_tbl.schema.field(c).type))
        tot_rows = dsum(float(sum(rows_in[r] for r in dict.fromkeys(order))))
        n_events = {"q1": tot_rows, "q2": tot_rows, "q5": tot_rows}.get(q, tot_rows * 50 / 4)   # persons+auctions are 4 of 50 events
        alg = dsum(algorithmic_bytes(q, rel, out_rows, extra))
        med, best = statistics.median(times), min(times)
        line = {"query": q, "n_gpus": world, "scale": scale, "rows_in": {k: dsum(float(v)) for k, v in rows_in.items()}, "rows_out": dsum(float(out_rows)),
                "device_ms_median": med, "device_ms_best": best, "rows_per_sec": tot_rows / (med * 1e-3), "stream_events_per_sec": n_events / (med * 1e-3),
                "algorithmic_bytes": alg, "achieved_gbs": alg / (med * 1e-3) / 1e9, "hbm_peak_gbs": peak * world,
                "frac_of_hbm_peak": alg / (med * 1e-3) / 1e9 / (peak * world), "e2e_ms": e2e_ms,
                "e2e_rows_per_sec": tot_rows / (e2e_ms * 1e-3) if e2e_ms else None, "kernels": prof,
                "cache": "L2 flushed (384 MB memset) before every repetition"}

        if args.check or (not args.no_cpu and world == 1 and sum(rows_in.values()) <= args.cpu_max_rows):
            import oracle
            cores = os.cpu_count() or 1
            src = [[rel[r]] for r in order]
            t = time.perf_counter()
            want = oracle.execute_plan(plans.QUERIES[q](min(cores, 16)), src, threads=min(cores, 16))
            cpu_s = time.perf_counter() - t
            line["cpu_oracle"] = {"ms": cpu_s * 1e3, "rows_per_sec": sum(rows_in[r] for r in dict.fromkeys(order)) / cpu_s, "threads": min(cores, 16)}
            if args.check and world == 1:
                oracle.assert_tables_equal(res_tbl, want)
                line["parity"] = "bit-exact vs oracle (sorted)"
        if rank == 0:
            print(json.dumps(line), file=_JSON_OUT, flush=True)
        ec.close()
        del resident, rel, out

    ctx.close()
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
